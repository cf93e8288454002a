"""debug aid: where do the fused and the split kernel path differ (first launches of a few shapes)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import mbe_testlib as T
from __graft_entry__ import load_package
pkg = load_package()
codec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
S = 64
frames = T.random_hard_frames(codec, S, 24, 5)
seeds = T.stream_seeds(S, 0x51)
for cuts in ([(0, 24)], [(0, 1), (1, 2), (2, 24)], [(0, 5), (5, 24)]):
    outs = []
    for path in (0, 1):
        dec = pkg.Decoder(max_streams=S, device=0)
        dec.set_kernel_path(path)
        dec.init_streams(0, S, seeds)
        parts = []
        states = []
        for a, b in cuts:
            parts.append(dec.process_frames(codec, np.ascontiguousarray(frames[:, a:b]), want_float=True)["pcmf"])
            states.append(dec.export_state(0, S).view(np.uint32).reshape(S, 3, -1))
        outs.append((np.concatenate(parts, axis=1), states))
        dec.close()
    a, b = outs[0][0], outs[1][0]
    bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
    print("cuts", cuts, "mismatching samples", len(bad))
    if len(bad):
        fr = sorted(set((int(x[0]), int(x[1])) for x in bad))
        print("  (stream, frame) pairs:", fr[:20], "... total", len(fr))
        s, f, n = bad[0]
        print("  first:", s, f, n, a[s, f, n:n + 6], b[s, f, n:n + 6])
    for i, (sa, sb) in enumerate(zip(outs[0][1], outs[1][1])):
        d = np.argwhere(sa != sb)
        if len(d):
            words = sorted(set((int(x[1]), int(x[2])) for x in d))
            print("  state after launch", i, "differs: streams", len(set(int(x[0]) for x in d)), "(struct, word):", words[:12], "... total", len(words))
# which frames: flags / L of the failing ones (last cuts layout, fused results)
dec = pkg.Decoder(max_streams=S, device=0); dec.set_kernel_path(0); dec.init_streams(0, S, seeds)
r = dec.process_frames(codec, frames, want_float=True)
a = r["pcmf"]
dec2 = pkg.Decoder(max_streams=S, device=0); dec2.set_kernel_path(1); dec2.init_streams(0, S, seeds)
b = dec2.process_frames(codec, frames, want_float=True)["pcmf"]
badf = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
fl = r["results"]["flags"]
print("failing frames", int(badf.sum()), "of", badf.size)
import collections
print("flags of failing:", collections.Counter(fl[badf].tolist()).most_common(6))
print("flags of passing:", collections.Counter(fl[~badf].tolist()).most_common(6))
for s in range(2):
    print("stream", s, "fail:", np.nonzero(badf[s])[0].tolist(), "flags:", [hex(int(x)) for x in fl[s]])
