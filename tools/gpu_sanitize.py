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
    # stage-level entry points on the same data
    fr1 = frames.reshape(S * F, -1).copy()
    dec.channel_step(codec, 0, frames=fr1)
    dec.channel_step(codec, 1, frames=fr1)
    d1 = np.zeros((S * F, pkg.PARAM_BITS[codec]), np.uint8)
    dec.channel_step(codec, 2, frames=fr1, bits=d1)
    if codec == 1:
        dec.channel_step(codec, 3, bits=d1)
    assert np.array_equal(d1.reshape(S, F, -1), r["bits"])
    st = dec.export_state(0, S)
    cur, prev, enh = st[:, 0].copy(), st[:, 1].copy(), st[:, 2].copy()
    dec.decode_parms(codec, d1[:S], cur, prev)
    dec.spectral_amp_enhance(cur)
    dec.adaptive_smoothing(cur, enh)
    print("codec", codec, "ok", int(np.abs(r["pcm"]).max()), int(np.abs(r2["pcm"]).max()))
words = rng.integers(0, 2, size=(64, 23), dtype=np.uint8)
rel = rng.integers(0, 256, size=(64, 23), dtype=np.uint8)
for code, ln in ((0, 23), (1, 15), (2, 15)):
    dec.ecc_blocks(code, words[:, :ln])
    dec.ecc_blocks(code, np.ascontiguousarray(np.stack([words[:, :ln], rel[:, :ln]], axis=-1)), soft=True)
cur = dec.export_state(0, 8)[:, 0].copy()
dec.synthesize_tone(cur, bits49=rng.integers(0, 2, size=(8, 49), dtype=np.uint8))
dec.synthesize_tone(cur, dstar_id=np.arange(8, dtype=np.int32) + 3)
dec.comfort_noise(dec.export_rng(0, 8))
print("stage entry points ok")
dec.close()
# the single-frame call of the shim (mbe_b200_single_frame): caller-owned state in, one frame, state out
import ctypes
dec = pkg.Decoder(max_streams=2, device=0)
lib = dec.lib
for codec in (0, 3):
    fr = T.random_hard_frames(codec, 1, 4, 31 + codec)
    trip = dec.export_state(0, 1)[0].copy()
    rng4 = dec.export_rng(0, 1)[0].copy()
    pcm = np.zeros(160, np.int16)
    res = np.zeros(1, pkg.RESULT_DTYPE)
    bits = np.zeros(pkg.PARAM_BITS[codec], np.uint8)
    for f in range(4):
        frame = np.ascontiguousarray(fr[0, f])
        rc = lib.mbe_b200_single_frame(dec.h, codec, 0, 0, frame.ctypes.data_as(ctypes.c_void_p), trip.ctypes.data_as(ctypes.c_void_p),
                                       rng4.ctypes.data_as(ctypes.c_void_p), pcm.ctypes.data_as(ctypes.c_void_p), None,
                                       res.ctypes.data_as(ctypes.c_void_p), bits.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0 and res["status"][0] >= 0
print("single-frame call ok")
dec.close()
