"""Host-side sharding of voice streams over the GPUs of one box.

Streams are independent (the reference is re-entrant per stream, include/mbelib-neo/mbelib.h:28-30), so the
multi-GPU story is a pure partition: rank r of `world` owns one contiguous block of global stream ids, runs the
same kernels on its own context and state pool, and nothing crosses GPUs on the data path.  The only
communication is bookkeeping (barrier, max-over-ranks of timings, optional gather of per-rank summaries), which
is why this module only needs torch.distributed as a thin optional dependency and works with the gloo backend
on CPU (tests/test_sharding_gloo.py) exactly as with NCCL on GPUs (bench.py).
"""
import numpy as np

SEED_BASE = 0xC0FFEE  # per-stream RNG convention: seed_s = SEED_BASE + global stream id


def shard_range(n_streams, rank, world):
    """Contiguous block of global stream ids [first, first + count) owned by `rank`; the first n_streams % world
    ranks get one extra stream."""
    if world < 1 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_streams, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def weak_shard(streams_per_gpu, rank):
    """Weak-scaling layout used by bench.py: every rank owns `streams_per_gpu` streams."""
    return rank * streams_per_gpu, streams_per_gpu


def stream_seeds(first, count, base=SEED_BASE):
    """mbe_setThreadRngSeed() argument of each stream, as a function of its GLOBAL id (uint32 wrap-around)."""
    return ((np.arange(count, dtype=np.uint64) + np.uint64(first) + np.uint64(base)) & np.uint64(0xffffffff)).astype(np.uint32)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, device=None):
    """All ranks' shard sizes (list of ints, rank order)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
