"""Many GPUs from one process (mbe_b200_pool_*): the stream range is sharded in contiguous blocks over
per-device contexts, nothing is exchanged between them.  On a one-GPU box the pool lists ordinal 0 several times (several
contexts on one GPU), which exercises the same sharding, offset and threading logic; with more GPUs visible it uses them
all.  Results must equal the single-context path bit for bit wherever a stream lands."""
import numpy as np
import pytest

import mbe_testlib as T
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_package()


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n_shards,codec", [(3, 3), (2, 0), (5, 2)])
def test_pool_equals_single_context(pkg, n_shards, codec):
    S, F = 100, 12                       # ragged: 100 streams do not divide evenly into 3 or 5 shards
    frames = T.random_hard_frames(codec, S, 2 * F, 0xA00 + codec)
    seeds = T.stream_seeds(S, 0xBEEF)
    one = pkg.Decoder(max_streams=S, device=0)
    one.init_streams(0, S, seeds)
    want_a = one.process_frames(codec, frames[:, :F], want_float=True)
    want_b = one.process_frames(codec, frames[:, F:], want_float=True)
    want_state = one.export_state(0, S)
    one.close()

    pool = pkg.Pool(S, devices=_devices(n_shards))
    sh = pool.shards()
    assert len(sh) == n_shards and sh[0][0] == 0 and sum(n for _, n in sh) == S
    assert all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(n_shards - 1))
    pool.init_streams(0, S, seeds)
    got_a = pool.process_frames(codec, frames[:, :F], want_float=True)
    # second half through two windows that straddle shard boundaries, the second one bit-packed
    cut = sh[0][1] + 3
    got_b1 = pool.process_frames(codec, frames[:cut, F:], first_stream=0, want_float=True)
    got_b2 = pool.process_frames(codec, pkg.pack_frames(codec, frames[cut:, F:]), first_stream=cut, want_float=True, packed=True)
    for k in ("pcm", "bits"):
        assert np.array_equal(got_a[k], want_a[k]), k
        assert np.array_equal(np.concatenate([got_b1[k], got_b2[k]]), want_b[k]), k
    assert np.array_equal(got_a["pcmf"].view(np.uint32), want_a["pcmf"].view(np.uint32))
    assert np.array_equal(np.concatenate([got_b1["results"], got_b2["results"]]), want_b["results"])
    assert np.array_equal(pool.export_state(0, S), want_state)
    # state moves between pools of different shapes (a stream changes device)
    other = pkg.Pool(S, devices=_devices(2 if n_shards != 2 else 3))
    other.import_state(want_state[10:90], first=10)
    assert np.array_equal(other.export_state(10, 80), want_state[10:90])
    other.close()
    pool.close()


def test_more_shards_than_divide_the_streams_and_a_channel_map(pkg):
    """9 streams over 8 shards: every shard owns at least one stream (2 1 1 1 1 1 1 1, base + remainder like
    sharding.shard_range), and a channel map set on the pool reaches every shard (it used to fail half-way on the empty
    trailing shards of a ceil() partition)."""
    codec, S, F = 3, 9, 6
    pool = pkg.Pool(S, devices=_devices(8))
    sh = pool.shards()
    assert [n for _, n in sh] == [2, 1, 1, 1, 1, 1, 1, 1] and [a for a, _ in sh] == [0, 2, 3, 4, 5, 6, 7, 8]
    rng = np.random.default_rng(99)
    pos = np.array(list(range(24)) + [24 + c for c in range(23)] + [48 + c for c in range(11)] + [72 + c for c in range(14)],
                   np.uint16)
    cmap = rng.permutation(pos)
    frames = T.random_hard_frames(codec, S, F, 0x909)
    mask = np.zeros(T.FRAME_BITS[codec], np.uint8)
    mask[pos] = 1
    frames &= mask
    air = np.packbits(frames[..., cmap], axis=-1, bitorder="big")
    seeds = T.stream_seeds(S, 17)
    pool.init_streams(0, S, seeds)
    want = pool.process_frames(codec, frames, want_float=True)
    pool.set_channel_map(codec, cmap)
    pool.init_streams(0, S, seeds)
    got = pool.process_frames(codec, air, want_float=True, packed=True)
    pool.set_channel_map(codec, None)
    for k in ("pcm", "bits", "results"):
        assert np.array_equal(got[k], want[k]), k
    cpu = T.run_cpu(T.load_oracle().mbo_run, codec, 0, frames, seeds)
    assert np.array_equal(want["pcm"], cpu["pcm"])
    pool.close()


def test_pool_range_errors(pkg):
    pool = pkg.Pool(16, devices=_devices(2))
    with pytest.raises(pkg.MbeB200Error):
        pool.init_streams(10, 7)
    with pytest.raises(pkg.MbeB200Error):
        pool.process_frames(3, np.zeros((17, 1, 96), np.uint8))
    pool.close()


def test_submit_and_wait_from_one_thread(pkg):
    """mbe_b200_submit_frames / mbe_b200_wait: one host thread queues a batch on several contexts, then collects them; a
    second submission on a busy context is refused; results equal the blocking call."""
    import torch
    codec, S, F = 3, 300, 8
    n_ctx = 3
    frames = [T.random_hard_frames(codec, S, F, 0xD00 + k) for k in range(n_ctx)]
    seeds = T.stream_seeds(S, 0x51)
    ref = pkg.Decoder(max_streams=S, device=0)
    want = []
    for k in range(n_ctx):
        ref.init_streams(0, S, seeds)
        want.append(ref.process_frames(codec, frames[k]))
    ref.close()
    decs = [pkg.Decoder(max_streams=S, device=k % torch.cuda.device_count()) for k in range(n_ctx)]
    pinned = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    ins, outs = [], []
    for k, d in enumerate(decs):
        d.init_streams(0, S, seeds)
        fr = pinned(frames[k].shape, torch.uint8)
        fr[...] = frames[k]
        out = dict(pcm=pinned((S, F, 160), torch.int16), bits=pinned((S, F, pkg.PARAM_BITS[codec]), torch.uint8),
                   results=pinned((S, F, 6), torch.int32).view(pkg.RESULT_DTYPE).reshape(S, F))
        d.submit_frames(codec, fr, out)
        ins.append(fr)
        outs.append(out)
    with pytest.raises(pkg.MbeB200Error):
        decs[0].submit_frames(codec, ins[0], outs[0])          # still pending
    for d in decs:
        d.wait()
    for k in range(n_ctx):
        for key in ("pcm", "bits", "results"):
            assert np.array_equal(outs[k][key], want[k][key]), (k, key)
    decs[0].wait()                                              # waiting twice is harmless
    again = decs[0].process_frames(codec, frames[0])            # and the context is usable again
    assert again["pcm"].shape == (S, F, 160)
    for d in decs:
        d.close()


def test_contexts_release_their_device_memory(pkg):
    """create / use / destroy in a loop: tables, state pool, staging buffers, streams and events all go back."""
    import torch
    codec, S, F = 3, 2048, 4
    frames = T.random_hard_frames(codec, S, F, 1)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for k in range(12):
        d = pkg.Decoder(max_streams=S, device=0)
        d.init_streams(0, S, None)
        d.process_frames(codec, frames, want_float=True)
        d.decode_frames(codec, frames.reshape(S * F, -1))
        d.close()
        p = pkg.Pool(S, devices=[0, 0])
        p.init_streams(0, S, None)
        p.process_frames(codec, frames)
        p.close()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < 64 << 20, "device memory not released: %d MB" % ((free0 - free1) >> 20)


def test_pinned_host_buffers_from_the_c_abi(pkg):
    """mbe_b200_host_alloc / host_register: a plain-C caller gets page-locked frame and PCM arrays without linking the CUDA
    runtime (the reference's "caller owns every buffer" contract, mbelib.h:28-30).  Results do not depend on
    where the host buffers live; a registered pageable array works the same and unregisters cleanly."""
    import ctypes
    lib = pkg.load_library()
    codec, S, F = 0, 300, 5
    frames = T.random_hard_frames(codec, S, F, 0xA110C)
    seeds = T.stream_seeds(S, 3)
    dec = pkg.Decoder(max_streams=S, device=0)
    dec.init_streams(0, S, seeds)
    want = dec.process_frames(codec, frames)
    h_fr = pkg.host_alloc(frames.shape, np.uint8)
    h_pcm = pkg.host_alloc((S, F, 160), np.int16)
    h_fr[...] = frames
    dec.init_streams(0, S, seeds)
    got = dec.process_frames(codec, h_fr, out_pcm=h_pcm)
    assert got["pcm"] is h_pcm and np.array_equal(h_pcm, want["pcm"])
    pkg.host_free(h_fr)
    pkg.host_free(h_pcm)
    own = np.zeros((S, F, 160), np.int16)
    assert lib.mbe_b200_host_register(own.ctypes.data_as(ctypes.c_void_p), own.nbytes) == 0
    dec.init_streams(0, S, seeds)
    dec.process_frames(codec, frames, out_pcm=own)
    assert lib.mbe_b200_host_unregister(own.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(own, want["pcm"])
    assert lib.mbe_b200_host_alloc(None, 16) != 0        # bad argument: error code, no crash
    dec.close()
