#!/usr/bin/env python
"""Per-stage cycle breakdown of the stream kernel (needs a library built with -DMBE_STAGE_TIMING=1, passed via
MBE_B200_LIB).  usage: MBE_B200_LIB=$PWD/build/var/timing.so python tools/gpu_stage_timing.py [codec] [streams]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
pkg = load_package()
codec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
F = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dec = pkg.Decoder(S, 0)
dec.init_streams(0, S, (np.arange(S) + 0xC0FFEE).astype(np.uint32))
fr = torch.randint(0, 2, (S, F, pkg.FRAME_BITS[codec]), dtype=torch.uint8, device="cuda")
pcm = torch.empty((S, F, 160), dtype=torch.int16, device="cuda")
import time
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.process_frames_dev(codec, 0, 0, S, F, fr.data_ptr(), pcm.data_ptr())
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    c = dec.debug_stage_cycles(reset=True).astype(np.float64)
print("launch %.3f ms for %d streams x %d frames = %.4g frames/s; a warp slot (148 SM x 28 warps) spends %.0f cycles per stream-launch at 1.965 GHz"
      % (wall * 1e3, S, F, S * F / wall, wall * 1.965e9 * 148 * 28 / S))
names = ["frame-top barrier", "front-end+decode+state machine", "enhance+synth_begin", "count barrier", "voiced bank (incl. barriers)",
         "unvoiced+handover", "output stores", "state store", "  bank: osc setup", "  bank: phase A", "  bank: interp",
         "  bank: wait A", "  bank: phase B", "  bank: wait B", "-", "-"]
tot = c[:14].sum()   # (the bank's inner timers take over the running clock: its time is the sum of the sub-stages)
for n, v in zip(names, c):
    print("%-34s %6.1f%%  %8.0f cycles/frame/warp" % (n, 100 * v / tot, v / (S * F)))
print("total %.0f cycles/frame/warp" % (tot / (S * F)))
