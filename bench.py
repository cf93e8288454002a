#!/usr/bin/env python
"""bench.py - decoded 20 ms frames/s of the batched IMBE/AMBE decode+synthesis path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (frame bits -> ECC -> parameter decode -> synthesis -> int16 PCM) over
one batch: BASELINE.json configs[1], AMBE+2 3600x2450, 65,536 streams x 50 synthetic random-bit frames per
GPU.  Streams are independent, so N GPUs run N disjoint stream shards with no collective on the data path
("scaling": "weak"); torch.distributed is used only for the barrier and the max-over-ranks of the timings.

Numbers in the JSON line:
  value        frames/s, whole job, inputs resident in HBM, timed with CUDA events on the launching stream.
  e2e          frames/s through the host-pointer C-ABI call (mbe_b200_process_frames): pinned HOST frame
               bits in, int16 PCM + results in HOST memory out, copies inside the timed region.
  roofline     HBM view of the stream kernel: algorithmic bytes per launch / mean launch time.
  roofline_fp32  the bound that actually applies (FP32 issue): algorithmic FLOP per launch / launch time
               against the non-fused FP32 issue peak (148 SM x 128 lanes x SM clock).
  cpu_baseline the reference's own CPU build (oracle/_ref, dev-release flags) on the box's host cores,
               bounded sample of the same workload.
`--impl reference` times only that CPU arm.  The CUDA library is mandatory for the default arm: there is no
CPU fallback in the product, and nothing under oracle/ is on the measured GPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CODEC_NAMES = {0: "imbe7200x4400", 1: "imbe7100x4400", 2: "ambe3600x2400", 3: "ambe3600x2450"}
FRAME_BITS = {0: 184, 1: 168, 2: 96, 3: 96}
# algorithmic FLOP per frame on iid random-bit frames (non-fused, mul = add = 1)
FLOP_PER_FRAME = {0: 84e3, 1: 84e3, 2: 66e3, 3: 61e3}
STATE_BYTES = 7828  # 3 x mbe_parms + 16 B RNG words per stream
# DRAM traffic of the stream kernel per frame, measured: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture (profiles/r01s_stream_kernel_ncu_details.txt: 172.4 MB + 367.4 MB for 16576 streams x 50
# frames of AMBE+2 hard-decision input) divided by the frames of that launch.  Only quoted for that workload.
NCU_DRAM_BYTES_PER_FRAME = {("ambe3600x2450", 0): (172.412928e6 + 367.429632e6) / (16576 * 50)}
# executed warp-instructions per frame of the same capture (smsp__inst_executed.sum / frames): the kernel is bound by
# instruction issue, so this x frames/s against 148 SM x 4 schedulers x SM clock is the utilisation that matters
NCU_WARP_INSTR_PER_FRAME = {("ambe3600x2450", 0): 4565407027.0 / (16576 * 50)}
RESULT_BYTES = 24
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region: NVML (nvidia_ml_py, one sample every few ms) when it
    loads, else `nvidia-smi` polling."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []      # (sm MHz, max MHz, watts, [active reason names])
        self.stop = threading.Event()
        self.th = None
        self.nvml = None
        self.handle = None
        self.source = "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        masks = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        self.rows.append((sm, mx, pw, [nm for nm, m in zip(self.NAMES, masks) if bits & m]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
            self.rows.append((float(parts[0]), float(parts[1]), float(parts[2]),
                              [nm for k, nm in enumerate(self.NAMES) if parts[3 + k].lower().startswith("active")]))

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.004 if self.nvml else 0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(r[0] for r in self.rows)
        reasons = [nm for nm in self.NAMES if any(nm in r[3] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in self.rows), "reasons": reasons,
                "power_w": round(max(r[2] for r in self.rows), 1), "samples": len(self.rows), "source": self.source}


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs next to its GPU: with eight ranks gathering
    PCM at once the host side is the limit, and remote-node staging makes it worse.  Returns a short
    description; does nothing when the box exposes no NUMA topology."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        if node < 0:
            return "numa: none exposed (%s)" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa: node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "numa: node %d, %d cpus (%s)" % (node, len(cpus), bdf)
    except Exception as e:
        return "numa: not bound (%s)" % type(e).__name__


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def channel_like_frames(codec, n_streams, n_frames, seed):
    """BASELINE.json configs[4]-shaped soft input: valid encoded frames (random bits for IMBE 7100, whose encoder the
    tests do not have), every channel bit flipped with p = 0.10, reliability 255 for unflipped and U[0,64) for flipped
    bits.  Returns uint8 [streams][frames][bits][2]."""
    import mbe_testlib as T
    rng = np.random.default_rng(seed)
    if codec == 1:
        hard = T.random_hard_frames(codec, n_streams, n_frames, 0x7100)
    else:
        enc = T.encode_imbe7200_frame if codec == 0 else T.encode_ambe_frame
        hard = np.zeros((n_streams, n_frames, FRAME_BITS[codec]), np.uint8)
        for b in range(n_streams):
            for f in range(n_frames):
                pb = rng.integers(0, 2, size=T.PARAM_BITS[codec], dtype=np.uint8)
                pb[0] = 0
                hard[b, f] = enc(pb).reshape(-1)
    return T.soften(hard, rng, flip_p=0.10)


def tone_unvoiced_frames(n_streams, n_frames, seed):
    """BASELINE.json configs[3]-shaped input (AMBE 3600x2400 only): frames built from parameter bits - a third tone / silence
    frames, a third unvoiced-only voice frames (FFT / WOLA / noise-generator path), a third ordinary voice frames - Golay
    encoded and PN scrambled into valid channel frames.  Returns uint8 [streams][frames][96]."""
    import mbe_testlib as T
    rng = np.random.default_rng(seed)
    frames = np.zeros((n_streams, n_frames, 96), np.uint8)
    for s in range(n_streams):
        for f in range(n_frames):
            p = rng.integers(0, 2, size=49, dtype=np.uint8)
            if (s + f // 5) % 3 == 0:
                p[0:6] = 1
            elif s % 2:
                p[38:42] = 0
            frames[s, f] = T.encode_ambe_frame(p).reshape(-1)
    return frames


def run_cpu_reference(codec, n_frames, steps, warmup, streams_per_thread, soft=0, soft_channel=False, tones=False):
    """Times the reference's CPU implementation (oracle/_ref dev-release build; the oracle port if the compiled
    reference is missing) on all host cores.  Returns (frames/s, info dict, seconds per step)."""
    import mbe_testlib as T
    cores = host_cores()
    lib, kind, fn = None, "reference", None
    if T.ref_available(fast=True):
        lib = T.load_ref(fast=True)
        fn = lib.ref_bench_run
        flavour = "oracle/_ref/libmberef_fast.so (unmodified reference, dev-release flags: SIMD + fast-math + LTO)"
    elif T.ref_available(fast=False):
        lib = T.load_ref(fast=False)
        fn = lib.ref_bench_run
        flavour = "oracle/_ref/libmberef.so (unmodified reference, Release flags)"
    else:
        fn = T.load_oracle().mbo_run
        kind = "port"
        flavour = "oracle/libmbe_oracle.so (C restatement)"
    if soft:
        streams_per_thread = max(1, streams_per_thread // 50)  # the reference's soft ECC is ~1 ms per IMBE frame
    S = min(cores * streams_per_thread, 65536)
    if tones:
        base = tone_unvoiced_frames(min(S, 128), n_frames, 0x2400)
        frames = np.ascontiguousarray(np.tile(base, ((S + len(base) - 1) // len(base), 1, 1))[:S])
    elif soft_channel:
        base = channel_like_frames(codec, min(S, 128), n_frames, 0x50F7)
        frames = np.ascontiguousarray(np.tile(base, ((S + len(base) - 1) // len(base), 1, 1, 1))[:S])
    elif soft:
        rng = np.random.default_rng(0x2450)
        frames = np.stack([T.random_hard_frames(codec, S, n_frames, 0x2450),
                           rng.integers(0, 256, size=(S, n_frames, FRAME_BITS[codec]), dtype=np.uint8)], axis=-1)
    else:
        frames = T.random_hard_frames(codec, S, n_frames, 0x2450)
    seeds = T.stream_seeds(S)
    pcm = np.zeros((S, n_frames, 160), np.int16)
    res = np.zeros((S, n_frames, 6), np.int32)
    times = []
    for it in range(warmup + steps):
        sec = fn(codec, int(bool(soft)), S, n_frames, T._ptr(frames), T._ptr(seeds), T._ptr(pcm), None, T._ptr(res), None, None, cores)
        if it >= warmup:
            times.append(sec)
    sec = float(np.mean(times))
    fps = S * n_frames / sec
    info = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "%d streams x %d frames per step (%d per thread), %s %s, one stream per task over %d pthreads, %s" % (
                S, n_frames, streams_per_thread, CODEC_NAMES[codec],
                "soft-decision (10% flipped bits)" if soft_channel else ("soft-decision" if soft else "hard-decision"), cores,
                flavour),
            "cpu_model": cpu_model(), "frames_per_s_per_core": fps / cores}
    return fps, info, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="ambe3600x2450", choices=list(CODEC_NAMES.values()))
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU")
    ap.add_argument("--frames", type=int, default=50, help="frames per stream per step")
    ap.add_argument("--soft", action="store_true", help="soft-decision input (bit + reliability per channel bit), random reliabilities")
    ap.add_argument("--soft-channel", action="store_true",
                    help="soft-decision input shaped like BASELINE.json configs[4]: valid encoded frames, every channel bit "
                         "flipped with p = 0.10, reliability 255 for unflipped and U[0,64) for flipped bits (128 distinct "
                         "streams tiled over the batch)")
    ap.add_argument("--tones-unvoiced", action="store_true",
                    help="BASELINE.json configs[3]-shaped input (forces --codec ambe3600x2400): tone, unvoiced-only and voice "
                         "frames from parameter bits, valid channel frames (128 distinct streams tiled over the batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.tones_unvoiced:
        args.codec = "ambe3600x2400"
    codec = {v: k for k, v in CODEC_NAMES.items()}[args.codec]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S, F = args.streams, args.frames
    soft = 1 if (args.soft or args.soft_channel) else 0
    workload = "%s %s-decision decode+synthesis, %d streams x %d synthetic %s frames per GPU" % (
        CODEC_NAMES[codec], "soft" if soft else "hard", S, F,
        "valid encoded, 10% flipped-bit" if args.soft_channel else ("tone / unvoiced-only / voice" if args.tones_unvoiced
                                                                      else "random-bit"))
    config = {"workload": workload, "codec": CODEC_NAMES[codec], "streams_per_gpu": S, "frames_per_stream": F,
              "sharding": "streams/%d, no collective" % world,
              "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % (S * F * (FRAME_BITS[codec] + 344) / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        warm = max(1, min(args.warmup, 3))
        fps, info, sec = run_cpu_reference(codec, F, args.steps, warm, streams_per_thread=1000, soft=soft,
                                           soft_channel=args.soft_channel, tones=args.tones_unvoiced)
        line = {"impl": "reference", "metric": "decoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "realtime_channels": fps / 50.0, "cpu_baseline": info,
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    pkg = load_package()
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "numa: single rank, not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    from importlib import import_module
    sharding = import_module("mbelib_neo_b200.sharding")
    first_global, _ = sharding.weak_shard(S, rank)      # this rank's block of global stream ids
    dec = pkg.Decoder(max_streams=S, device=local_rank)
    dec.init_streams(0, S, sharding.stream_seeds(first_global, S))

    fb = FRAME_BITS[codec]
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x2450 + rank)
    d_frames = torch.randint(0, 2, (S, F, fb), dtype=torch.uint8, device=dev, generator=gen)
    if args.tones_unvoiced:
        B = 128
        base = torch.from_numpy(tone_unvoiced_frames(B, F, 0x2400 + rank)).to(dev)
        d_frames = base.repeat((S + B - 1) // B, 1, 1)[:S].contiguous()
    elif args.soft_channel:
        B = 128
        base = torch.from_numpy(channel_like_frames(codec, B, F, 0x50F7 + rank)).to(dev)
        d_frames = base.repeat((S + B - 1) // B, 1, 1, 1)[:S].contiguous()
    elif soft:  # mbe_soft_bit {bit, reliability} pairs
        rel = torch.randint(0, 256, (S, F, fb), dtype=torch.uint8, device=dev, generator=gen)
        d_frames = torch.stack((d_frames, rel), dim=-1).contiguous()
    d_pcm = torch.empty((S, F, 160), dtype=torch.int16, device=dev)
    d_res = torch.empty((S, F, 6), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)

    # Short launches (a real-time server decodes ONE 20 ms frame per stream per launch) must not see the same frames
    # launch after launch: a stream fed the same frame twice has a perfectly stable pitch, which sends every low harmonic
    # through the phase-interpolated (one cosf per sample) path and times a workload nobody has.  Rotate through enough
    # different frame sets to cover 50 frames per stream.
    rot_sets = [d_frames]
    if F < 50 and not (args.soft_channel or args.tones_unvoiced):
        for k in range(1, min(50 // F, 25)):
            x = torch.randint(0, 2, (S, F, fb), dtype=torch.uint8, device=dev, generator=gen)
            if soft:
                x = torch.stack((x, torch.randint(0, 256, (S, F, fb), dtype=torch.uint8, device=dev, generator=gen)), dim=-1).contiguous()
            rot_sets.append(x)
    step_no = [0]

    def step_dev():
        fr = rot_sets[step_no[0] % len(rot_sets)]
        step_no[0] += 1
        dec.process_frames_dev(codec, soft, 0, S, F, fr.data_ptr(), d_pcm.data_ptr(), 0, d_res.data_ptr(), 0,
                               stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, device=dev)

    # ---- device-resident arm ----
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_dev()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = dec.launches
    with ClockSampler(local_rank) as clk:
        barrier()
        with torch.cuda.stream(stream):
            evs[0].record(stream)
            for k in range(args.steps):
                step_dev()
                evs[k + 1].record(stream)
        stream.synchronize()
        barrier()
    launches = dec.launches - l0
    total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]))
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    kern_ms = float(np.mean(per_launch_ms))
    value = world * S * F * args.steps / (total_ms * 1e-3)
    clocks = clk.summary()

    # sanity: the PCM that came out is not silence and statuses are valid
    chk = d_pcm[:64].abs().max().item()
    st_min = int(d_res[:, :, 0].min().item())
    if chk == 0 or st_min < 0:
        raise SystemExit("bench.py: device path produced silence or error statuses (max |pcm| %d, min status %d)" % (chk, st_min))

    # ---- end-to-end arm: host frames in, host PCM + results out, through the host-pointer C-ABI call ----
    e2e = None
    if not args.no_e2e:
        h_sets = []
        for x in rot_sets:       # (short launches rotate through different frame sets here too)
            hx = torch.empty(tuple(x.shape), dtype=torch.uint8, pin_memory=True)
            hx.copy_(x)
            h_sets.append(hx)
        h_frames = h_sets[0]
        h_pcm = torch.empty((S, F, 160), dtype=torch.int16, pin_memory=True)
        h_res = torch.empty((S, F, 6), dtype=torch.int32, pin_memory=True)
        np_frames, np_pcm = h_frames.numpy(), h_pcm.numpy()
        np_sets = [x.numpy() for x in h_sets]
        np_res = h_res.numpy().view(pkg.RESULT_DTYPE).reshape(S, F)
        lib, h = dec.lib, dec.h
        import ctypes
        host_step = [0]

        def step_host():
            cur_frames = np_sets[host_step[0] % len(np_sets)]
            host_step[0] += 1
            rc = lib.mbe_b200_process_frames(h, codec, soft, 0, S, F, cur_frames.ctypes.data_as(ctypes.c_void_p),
                                             np_pcm.ctypes.data_as(ctypes.c_void_p), None,
                                             np_res.ctypes.data_as(ctypes.c_void_p), None)
            if rc != 0:
                raise SystemExit("mbe_b200_process_frames failed: %s" % lib.mbe_b200_last_error(h).decode())

        for _ in range(max(1, min(args.warmup, 3))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()          # returns when PCM + results are in host memory
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_s = max_over_ranks(t1 - t0)
        barrier()
        e2e = {"value": world * S * F * args.steps / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": int(S * F * fb * (2 if soft else 1)), "d2h_bytes_per_step": int(S * F * (320 + RESULT_BYTES)),
               "ms_per_step": e2e_s * 1e3 / args.steps, "host_binding": numa_note,
               "note": "mbe_b200_process_frames: pinned host bits in, int16 PCM + results to pinned host memory, "
                       "wall clock around the blocking calls, max over ranks"}
        if float(np.abs(np_pcm[:64]).max()) == 0:
            raise SystemExit("bench.py: e2e path produced silence")
        if not soft and len(rot_sets) == 1:
            # the same call with bit-packed channel frames: 8x less host->device traffic
            h_packed = torch.from_numpy(pkg.pack_frames(codec, np_frames)).pin_memory()
            np_packed = h_packed.numpy()

            def step_packed():
                rc = lib.mbe_b200_process_frames_packed(h, codec, 0, S, F, np_packed.ctypes.data_as(ctypes.c_void_p),
                                                        np_pcm.ctypes.data_as(ctypes.c_void_p), None,
                                                        np_res.ctypes.data_as(ctypes.c_void_p), None)
                if rc != 0:
                    raise SystemExit("mbe_b200_process_frames_packed failed: %s" % lib.mbe_b200_last_error(h).decode())

            step_packed()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_packed()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            p_s = max_over_ranks(t1 - t0)
            barrier()
            e2e["packed_input"] = {"value": world * S * F * args.steps / p_s, "unit": "frames/s",
                                   "h2d_bytes_per_step": int(np_packed.size),
                                   "note": "mbe_b200_process_frames_packed: hard bits packed eight per byte"}

    # ---- roofline of the stream kernel ----
    hbm_peak, peak_src = measured_peaks()
    alg_bytes = S * F * (fb * (2 if soft else 1) + 320 + RESULT_BYTES) + 2 * S * STATE_BYTES
    hbm_ach = alg_bytes / (kern_ms * 1e-3) / 1e9
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_issue_peak = 148 * 128 * sm_mhz * 1e6 / 1e12  # TFLOP/s, one non-fused op per lane per clock
    flops = S * F * FLOP_PER_FRAME[codec]
    fp32_ach = flops / (kern_ms * 1e-3) / 1e12
    per_frame_dram = NCU_DRAM_BYTES_PER_FRAME.get((CODEC_NAMES[codec], soft))
    roofline = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "traffic": (per_frame_dram * S * F) if per_frame_dram else None,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per frame (profiles/r01s_*) x frames per launch"
                if per_frame_dram else None,
                "peak_source": peak_src, "kernel": "mbe_stream_kernel",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern_ms,
                "note": "the kernel is FP32-issue bound, not HBM bound (no dense contraction, ~1 KB/frame); see roofline_fp32"}
    roofline_fp32 = {"bound": "fp32-issue", "achieved": fp32_ach, "peak": fp32_issue_peak, "unit": "TFLOP/s",
                     "frac": fp32_ach / fp32_issue_peak, "flop_per_frame": FLOP_PER_FRAME[codec],
                     "peak_source": "148 SM x 128 FP32 lanes x %.0f MHz sampled SM clock, non-fused (mul and add issue "
                                    "separately because parity forbids FMA contraction); FMA-counted peak is 2x" % sm_mhz}

    wipf = NCU_WARP_INSTR_PER_FRAME.get((CODEC_NAMES[codec], soft))
    if wipf:
        issue_peak = 148 * 4 * sm_mhz * 1e6   # warp-instructions per second, one per scheduler per clock
        issue_ach = wipf * S * F / (kern_ms * 1e-3)
        roofline_fp32["issue_slots"] = {
            "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instr/s", "frac": issue_ach / issue_peak,
            "warp_instr_per_frame": wipf,
            "source": "ncu smsp__inst_executed.sum per frame (profiles/r01s_*) x frames per launch / event-timed launch"}
    line = {"metric": "decoded frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "realtime_channels": value / 50.0, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fp32": roofline_fp32}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, info, _ = run_cpu_reference(codec, F, 2, 1, streams_per_thread=1000, soft=soft, soft_channel=args.soft_channel,
                                       tones=args.tones_unvoiced)
        line["cpu_baseline"] = info
    dec.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
