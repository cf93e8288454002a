"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the unmodified reference's CPU
build from oracle/_ref (or the oracle port when it is absent) and prints one JSON line with the keys the driver reads;
the default arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from __graft_entry__ import ROOT


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "decoded frames/sec" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "ambe3600x2450" in line["config"]["workload"] and "65536 streams x 50" in line["config"]["workload"]


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
