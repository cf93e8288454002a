#!/usr/bin/env python
"""One bench line per BASELINE.json config on this GPU box (N = 1): configs[1] AMBE+2 hard (the headline, bench.py's default),
configs[2] IMBE 7200x4400 hard at its per-GPU share of 1M streams, configs[3] AMBE 3600x2400 tone / unvoiced-heavy frames,
configs[4] the mixed-codec soft-decision workload (a third of the streams per codec: the three launches run back to back,
so the mix decodes total frames / total time = 3 / sum(1 / rate_i)), each with the CPU reference arm beside it.
usage: python tools/bench_configs.py [--steps K] [--quick]  -> JSON lines on stdout"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = "3"
quick = "--quick" in sys.argv
if "--steps" in sys.argv:
    steps = sys.argv[sys.argv.index("--steps") + 1]


def run(flags, reference=False):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2" if reference else steps, "--warmup", "1" if reference else "3"]
    cmd += ["--impl", "reference"] if reference else ["--no-cpu-baseline"]
    out = subprocess.run(cmd + flags, capture_output=True, text=True, timeout=1200)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        raise SystemExit("bench.py %s failed:\n%s" % (" ".join(flags), out.stderr[-2000:]))
    return json.loads(lines[-1])


def brief(d):
    e2e = d.get("e2e") or {}
    return {"frames_per_s": d["value"], "ms_per_step": d["ms_per_step"], "e2e_frames_per_s": e2e.get("value"),
            "workload": d["config"]["workload"]}


per_gpu = "16384" if quick else "131072"
cases = [("configs[1]", []),
         ("configs[2]", ["--codec", "imbe7200x4400", "--streams", per_gpu]),
         ("configs[3]", ["--tones-unvoiced", "--streams", "32768" if not quick else "8192"])]
for name, flags in cases:
    g, c = run(flags), run(flags, reference=True)
    print(json.dumps({"config": name, "gpu": brief(g), "cpu_reference": {"frames_per_s": c["value"], "cores": c["cpu_baseline"]["cores"],
                                                                          "sample": c["cpu_baseline"]["sample"]}}), flush=True)
third = str(int(per_gpu) // 3)
gpu, cpu = [], []
for codec in ("imbe7200x4400", "imbe7100x4400", "ambe3600x2450"):
    flags = ["--codec", codec, "--soft-channel", "--streams", third]
    gpu.append(run(flags))
    cpu.append(run(flags, reference=True))
mix = lambda rs: 3.0 / sum(1.0 / r["value"] for r in rs)
print(json.dumps({"config": "configs[4]", "gpu": {"frames_per_s": mix(gpu), "per_codec": {r["config"]["codec"]: r["value"] for r in gpu},
                                                   "e2e_frames_per_s": 3.0 / sum(1.0 / r["e2e"]["value"] for r in gpu),
                                                   "workload": "mixed-codec soft-decision, %s streams per codec x 50 frames, valid frames with 10%% flipped bits" % third},
                  "cpu_reference": {"frames_per_s": mix(cpu), "per_codec": {r["config"]["codec"]: r["value"] for r in cpu},
                                    "cores": cpu[0]["cpu_baseline"]["cores"]}}), flush=True)
