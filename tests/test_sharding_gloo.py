"""The N > 1 path on CPU: two ranks over the gloo backend partition the streams exactly as bench.py does over
NCCL (no data-path collective), each decodes its shard, and the union equals the single-process result.  The CPU
oracle stands in for the GPU kernels here (test infrastructure only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mbe_testlib as T
from __graft_entry__ import ROOT, load_package

WORLD = 2
N_STREAMS, N_FRAMES, CODEC = 37, 6, 3     # odd stream count: ragged shards


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    pkg = load_package()
    from importlib import import_module
    sh = import_module("mbelib_neo_b200.sharding")
    first, count = sh.shard_range(N_STREAMS, rank, WORLD)
    counts = sh.gather_counts(count)
    assert sum(counts) == N_STREAMS and counts[rank] == count
    frames = T.random_hard_frames(CODEC, N_STREAMS, N_FRAMES, 4321)[first:first + count]
    seeds = sh.stream_seeds(first, count)
    out = T.run_cpu(T.load_oracle().mbo_run, CODEC, 0, frames, seeds)
    t = sh.max_over_ranks(1.0 + rank)            # the timing reduction bench.py uses
    assert t == float(WORLD)
    dist.barrier()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), first=first, count=count, pcm=out["pcm"], bits=out["bits"])
    dist.destroy_process_group()


def test_two_rank_partition_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    pkg = load_package()
    from importlib import import_module
    sh = import_module("mbelib_neo_b200.sharding")
    frames = T.random_hard_frames(CODEC, N_STREAMS, N_FRAMES, 4321)
    whole = T.run_cpu(T.load_oracle().mbo_run, CODEC, 0, frames, sh.stream_seeds(0, N_STREAMS))
    covered = np.zeros(N_STREAMS, bool)
    for r in range(WORLD):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        f, c = int(z["first"]), int(z["count"])
        assert not covered[f:f + c].any()
        covered[f:f + c] = True
        assert np.array_equal(z["pcm"], whole["pcm"][f:f + c])
        assert np.array_equal(z["bits"], whole["bits"][f:f + c])
    assert covered.all()


def test_shard_ranges_are_a_partition():
    pkg = load_package()
    from importlib import import_module
    sh = import_module("mbelib_neo_b200.sharding")
    for n in (0, 1, 7, 8, 65536, 1048576 + 3):
        for world in (1, 2, 4, 8):
            nxt = 0
            for r in range(world):
                f, c = sh.shard_range(n, r, world)
                assert f == nxt and c >= 0
                nxt = f + c
            assert nxt == n
    assert list(sh.stream_seeds(0xffffffff - 0xC0FFEE, 2)) == [0xffffffff, 0]
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)
