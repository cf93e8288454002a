"""Debug helper (run under gpurun): first field of the exported state that differs from the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mbe_testlib as T
from __graft_entry__ import load_package
pkg = load_package()
codec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = 256
dec = pkg.Decoder(S)
frames = T.random_hard_frames(codec, S, 50, 100 + codec)
seeds = T.stream_seeds(S)
for nf in (1, 2, 3, 5):
    dec.init_streams(0, S, seeds)
    got = dec.process_frames(codec, frames[:, :nf], want_float=True)
    want = T.run_cpu(T.load_oracle().mbo_run, codec, 0, frames[:, :nf], seeds)
    st = dec.export_state(0, S)
    bad = np.where((st != want["state"]).any(axis=(1, 2)))[0]
    fbad = (got["pcmf"].view(np.uint32) != want["pcmf"].view(np.uint32))
    print("frames", nf, "streams with state diff", len(bad), "float pcm diffs", fbad.sum(), "of", fbad.size)
    if len(bad):
        cnt = {}
        for s in bad:
            for which in range(3):
                a, b = T.parms_view(st[s, which]), T.parms_view(want["state"][s, which])
                for k in a:
                    x, y = np.atleast_1d(a[k]), np.atleast_1d(b[k])
                    if x.dtype == np.float32:
                        ne = x.view(np.uint32) != y.view(np.uint32)
                    else:
                        ne = x != y
                    if ne.any():
                        cnt[(which, k)] = cnt.get((which, k), 0) + 1
        print("  field diff counts:", sorted(cnt.items(), key=lambda kv: -kv[1])[:12])
        s = bad[0]
        for which in range(3):
            a, b = T.parms_view(st[s, which]), T.parms_view(want["state"][s, which])
            for k in a:
                x, y = np.atleast_1d(a[k]), np.atleast_1d(b[k])
                ne = (x.view(np.uint32) != y.view(np.uint32)) if x.dtype == np.float32 else (x != y)
                if ne.any():
                    i = np.where(ne)[0]
                    print("  stream", s, "struct", which, k, "idx", i[:6], x[i[:3]], y[i[:3]], "L", a["L"], "flags", want["results"][s, :, 5])
        break
