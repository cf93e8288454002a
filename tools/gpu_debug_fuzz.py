import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import mbe_testlib as T
from __graft_entry__ import load_package
pkg = load_package()
codec = 3
rng = np.random.default_rng(0xF022 + codec)
S, F = 4096, 24
fb, pb = T.FRAME_BITS[codec], T.PARAM_BITS[codec]
enc = T.encode_ambe_frame
P = 96
pool = np.zeros((P, fb), np.uint8)
for i in range(P):
    p = rng.integers(0, 2, size=pb, dtype=np.uint8)
    if i % 6 == 0:
        p[0:6] = 1
        if i % 12 == 0:
            p[45:49] = 0
    pool[i] = enc(p).reshape(-1)
kind = rng.choice(6, size=(S, F), p=[0.22, 0.30, 0.22, 0.18, 0.06, 0.02])
pick = rng.integers(0, P, size=(S, F))
frames = np.zeros((S, F, fb), np.uint8)
last = pool[pick[:, 0]]
for f in range(F):
    k = kind[:, f]
    fresh = pool[pick[:, f]]
    cur = np.where((k == 2)[:, None], last, fresh)
    last = np.where(((k == 1) | (k == 2) | (k == 3))[:, None], cur, last)
    noisy = cur ^ (rng.random((S, fb)) < rng.uniform(0.01, 0.08, size=(S, 1))).astype(np.uint8)
    rnd = rng.integers(0, 2, size=(S, fb), dtype=np.uint8)
    out = np.where((k == 0)[:, None], rnd, np.where((k == 3)[:, None], noisy, cur))
    sig = pool[(pick[:, f] // 6) * 6 % P]
    out = np.where((k == 4)[:, None], sig, out)
    bad = out.copy(); bad[:, 7] = 2
    frames[:, f] = np.where((k == 5)[:, None], bad, out)
seeds = T.stream_seeds(S, 0xF0 + codec)
want = T.run_cpu(T.load_oracle().mbo_run, codec, 0, frames, seeds, n_threads=16)
ref = T.load_ref()
if ref is not None:
    w2 = T.run_cpu(ref.ref_bench_run, codec, 0, frames, seeds, n_threads=16)
    print("oracle vs compiled reference: state equal", np.array_equal(want["state"], w2["state"]), "pcm equal", np.array_equal(want["pcm"], w2["pcm"]))
for path in (0, 1):
    dec = pkg.Decoder(max_streams=S, device=0); dec.set_kernel_path(path); dec.init_streams(0, S, seeds)
    got = dec.process_frames(codec, frames, want_float=True)
    st = dec.export_state(0, S).view(np.uint32).reshape(S, 3, -1)
    ws = want["state"].view(np.uint32).reshape(S, 3, -1)
    d = np.argwhere(st != ws)
    print("path", path, "pcm equal", np.array_equal(got["pcm"], want["pcm"]), "state diff entries", len(d), "streams", len(set(d[:, 0].tolist())))
    if len(d):
        import collections
        print("  (struct, word) histogram:", collections.Counter((int(a), int(b)) for _, a, b in d).most_common(12))
        s0 = int(d[0][0])
        fl = got["results"]["flags"][s0]
        print("  stream", s0, "flags", [hex(int(x)) for x in fl], "status", got["results"]["status"][s0].tolist())
        for _, a, b in d[d[:, 0] == s0][:6]:
            print("   struct", a, "word", b, "gpu", hex(int(st[s0, a, b])), "oracle", hex(int(ws[s0, a, b])))
    dec.close()
