"""Where does the host-pointer call's time go?  Times mbe_b200_process_frames on the bench workload with different
output sets (PCM + results, results only) and with the pipeline shape taken from the environment, next to the
device-resident launch.  Prints one line per case.

    MBE_B200_KSTREAMS=4 MBE_B200_TAPER=18 MBE_B200_CHUNKS=32 python tools/gpu_e2e_probe.py
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402


def main():
    pkg = load_package()
    S, F, fb, codec = 65536, 50, 96, 3
    if len(sys.argv) > 2:      # python tools/gpu_e2e_probe.py <codec 0..3> <streams>
        codec, S = int(sys.argv[1]), int(sys.argv[2])
        fb = {0: 184, 1: 168, 2: 96, 3: 96}[codec]
    dev = torch.device("cuda", 0)
    dec = pkg.Decoder(max_streams=S, device=0)
    dec.init_streams(0, S, np.arange(S, dtype=np.uint32) + 0xC0FFEE)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x2450)
    d_frames = torch.randint(0, 2, (S, F, fb), dtype=torch.uint8, device=dev, generator=gen)
    h_frames = torch.empty((S, F, fb), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    h_pcm = torch.empty((S, F, 160), dtype=torch.int16, pin_memory=True)
    h_res = torch.empty((S, F, 6), dtype=torch.int32, pin_memory=True)
    d_pcm = torch.empty((S, F, 160), dtype=torch.int16, device=dev)
    d_res = torch.empty((S, F, 6), dtype=torch.int32, device=dev)
    lib, h = dec.lib, dec.h
    pf, pp, pr = (ctypes.c_void_p(x.data_ptr()) for x in (h_frames, h_pcm, h_res))

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps

    def host(pcm, res):
        rc = lib.mbe_b200_process_frames(h, codec, 0, 0, S, F, pf, pcm, None, res, None)
        assert rc == 0, lib.mbe_b200_last_error(h)

    stream = torch.cuda.Stream(device=dev)
    t_dev = timed(lambda: dec.process_frames_dev(codec, 0, 0, S, F, d_frames.data_ptr(), d_pcm.data_ptr(), 0, d_res.data_ptr(), 0,
                                                 stream.cuda_stream))
    t_full = timed(lambda: host(pp, pr))
    t_res = timed(lambda: host(None, pr))
    t_d2h = timed(lambda: h_pcm.copy_(d_pcm, non_blocking=True))
    print("shape k=%s taper=%s chunks=%s | device-resident %.2f ms | host call pcm+results %.2f ms | results only %.2f ms | "
          "plain PCM d2h %.2f ms" % (os.environ.get("MBE_B200_KSTREAMS", "-"), os.environ.get("MBE_B200_TAPER", "-"),
                                         os.environ.get("MBE_B200_CHUNKS", "-"), t_dev, t_full, t_res, t_d2h))


if __name__ == "__main__":
    main()
