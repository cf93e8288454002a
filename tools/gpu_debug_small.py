import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import mbe_testlib as T
from __graft_entry__ import load_package
pkg = load_package()
codec, S, F = 3, 9, 6
frames = T.random_hard_frames(codec, S, F, 0x909)
seeds = T.stream_seeds(S, 17)
cpu = T.run_cpu(T.load_oracle().mbo_run, codec, 0, frames, seeds)
nbad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    pool = pkg.Pool(S, devices=[0] * 8)
    pool.init_streams(0, S, seeds)
    got = pool.process_frames(codec, frames, want_float=True)
    bad = np.argwhere((got["pcm"] != cpu["pcm"]).any(axis=2)).tolist()
    st = pool.export_state(0, S)
    if bad:
        nbad += 1
        s, f = bad[0]
        d = np.nonzero(got["pcm"][s, f] != cpu["pcm"][s, f])[0]
        print("iter", it, "bad frames", bad, "first bad samples", d[:8].tolist(), got["pcm"][s, f, d[:4]].tolist(), cpu["pcm"][s, f, d[:4]].tolist(),
              "state equal", np.array_equal(st, cpu["state"]), "bits eq", np.array_equal(got["bits"], cpu["bits"]))
    pool.close()
print("bad iterations", nbad)
