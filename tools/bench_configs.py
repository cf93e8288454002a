#!/usr/bin/env python
"""One bench line per BASELINE.json config on this GPU box (N = 1), each at the config's full size and with the CPU reference arm
(bench.py's bounded cpu_baseline sample) beside it: configs[1] AMBE+2 hard 65,536 x 50, configs[2] IMBE 7200x4400 hard
1,048,576 x 50 (the headline, bench.py's default), configs[3] AMBE 3600x2400 tone / unvoiced-only / voice frames 262,144 x 50,
configs[4] mixed-codec soft decision at 10 % flipped bits, 1,048,576 streams (a third per codec).
usage: python tools/bench_configs.py [--steps K]  -> JSON lines on stdout"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = sys.argv[sys.argv.index("--steps") + 1] if "--steps" in sys.argv else "3"
for cfg in (1, 2, 3, 4):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", str(cfg), "--steps", steps, "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        print(json.dumps({"config": "configs[%d]" % cfg, "error": out.stderr[-600:]}), flush=True)
        continue
    d = json.loads(lines[-1])
    e2e = d.get("e2e") or {}
    cpu = d.get("cpu_baseline") or {}
    r = d.get("roofline") or {}
    print(json.dumps({"config": "configs[%d]" % cfg, "workload": d["config"]["workload"], "frames_per_s": d["value"],
                      "ms_per_step": d["ms_per_step"], "e2e_frames_per_s": e2e.get("value"), "e2e_frac_of_link": e2e.get("frac_of_link"),
                      "kernel_path": d.get("kernel_path"), "gpu_launches": d.get("gpu_launches"),
                      "roofline": {k: r.get(k) for k in ("kernel", "achieved", "peak", "frac", "share_of_step")},
                      "kernels": d.get("kernels"), "clocks": d.get("clocks"),
                      "cpu_reference": {"frames_per_s": cpu.get("value"), "cores": cpu.get("cores"), "sample": cpu.get("sample")}}), flush=True)
