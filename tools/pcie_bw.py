"""Host link probe for the e2e figure: pinned-memory device->host and host->device copy bandwidth of this box, each alone
and both at once (the shape of the e2e step: 1.13 GB of PCM + results out, 0.31 GB of channel bits in).

    python tools/pcie_bw.py            # prints one JSON line
"""
import json

import torch


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    out_bytes, in_bytes = 1127219200, 314572800  # bench.py's d2h / h2d bytes per step (65 536 streams x 50 frames)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    s2 = torch.cuda.Stream()
    d2h = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    h2d = timed(lambda: d_in.copy_(h_in, non_blocking=True))

    def both():
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(s2):
            s2.wait_event(ev)
            d_in.copy_(h_in, non_blocking=True)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)

    duplex = timed(both)
    print(json.dumps({
        "d2h_GBps": out_bytes / d2h / 1e6, "d2h_ms": d2h,
        "h2d_GBps": in_bytes / h2d / 1e6, "h2d_ms": h2d,
        "duplex_ms": duplex, "duplex_d2h_GBps": out_bytes / duplex / 1e6,
        "note": "pinned host memory, best of 5, CUDA events; duplex = the 1.13 GB d2h and 0.31 GB h2d of one bench step issued together",
    }))


if __name__ == "__main__":
    main()
