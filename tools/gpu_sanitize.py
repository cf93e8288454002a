#!/usr/bin/env python
"""Small batches of every codec through the stream kernel; run under compute-sanitizer:
   compute-sanitizer --tool racecheck|memcheck|synccheck python tools/gpu_sanitize.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mbe_testlib as T
from __graft_entry__ import load_package
pkg = load_package()
S, F = 45, 6          # 4 blocks of 14 streams, last one ragged
dec = pkg.Decoder(max_streams=S, device=0)
rng = np.random.default_rng(5)
for codec in (3, 0, 2, 1):
    frames = T.random_hard_frames(codec, S, F, 900 + codec)
    dec.init_streams(0, S, T.stream_seeds(S))
    r = dec.process_frames(codec, frames, want_float=True)
    soft = T.soften(frames, rng, flip_p=0.05)
    dec.init_streams(0, S, T.stream_seeds(S))
    r2 = dec.process_frames(codec, soft, soft=True)
    # random reliabilities: every row runs the full soft-decision search
    soft[..., 1] = rng.integers(0, 256, size=soft[..., 1].shape)
    dec.init_streams(0, S, T.stream_seeds(S))
    r2b = dec.process_frames(codec, soft, soft=True)
    bits, res = dec.decode_frames(codec, soft.reshape(S * F, -1), soft=True)
    assert np.array_equal(bits.reshape(S, F, -1), r2b["bits"])
    dec.init_streams(0, S, T.stream_seeds(S))
    r3 = dec.process_frames_packed(codec, pkg.pack_frames(codec, frames))
    assert np.array_equal(r["pcm"], r3["pcm"])
    print("codec", codec, "ok", int(np.abs(r["pcm"]).max()), int(np.abs(r2["pcm"]).max()))
dec.close()
