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
F = 50
dec = pkg.Decoder(S, 0)
dec.init_streams(0, S, (np.arange(S) + 0xC0FFEE).astype(np.uint32))
fr = torch.randint(0, 2, (S, F, pkg.FRAME_BITS[codec]), dtype=torch.uint8, device="cuda")
pcm = torch.empty((S, F, 160), dtype=torch.int16, device="cuda")
for it in range(2):
    dec.process_frames_dev(codec, 0, 0, S, F, fr.data_ptr(), pcm.data_ptr())
    torch.cuda.synchronize()
    c = dec.debug_stage_cycles(reset=True).astype(np.float64)
names = ["frame-top barrier", "front-end+decode+state machine", "enhance+synth_begin", "count barrier", "voiced bank (incl. barriers)",
         "unvoiced+handover", "output stores", "state store", "  bank: osc setup", "  bank: phase A", "  bank: interp",
         "  bank: wait A", "  bank: phase B", "  bank: wait B", "-", "-"]
tot = c[:8].sum()
for n, v in zip(names, c):
    print("%-34s %6.1f%%  %8.0f cycles/frame/warp" % (n, 100 * v / tot, v / (S * F)))
print("total %.0f cycles/frame/warp" % (tot / (S * F)))
