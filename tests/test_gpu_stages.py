"""SURVEY 8(a) rows P1-P5 on their own: parameter dequantisation (mbe_decode<Codec>Parms), spectral amplitude
enhancement and adaptive smoothing run as single stages on caller-held mbe_parms blobs (mbe_b200_decode_parms,
mbe_b200_spectral_amp_enhance, mbe_b200_adaptive_smoothing) against the oracle's per-function API, chained over
several frames so that the prediction state (prev_mp, mutated by the decoders: trap T3) is exercised.  Bit-exact:
the blobs are compared byte for byte."""
import ctypes

import numpy as np
import pytest

import mbe_testlib as T
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def dec():
    d = load_package().Decoder(max_streams=4, device=0)
    yield d
    d.close()


def _fresh(n):
    """n x {cur, prev, enh} as mbe_initMbeParms leaves them (from the device: init_streams + export_state)."""
    d = load_package().Decoder(max_streams=n, device=0)
    d.init_streams(0, n, None)
    st = d.export_state(0, n)
    d.close()
    return st


@pytest.mark.parametrize("codec", [0, 2, 3])
def test_decode_parms_chain(dec, codec):
    o = T.load_oracle()
    n, frames = 96, 6
    pb = T.PARAM_BITS[codec]
    rng = np.random.default_rng(0x9A0 + codec)
    st = _fresh(n)
    g_cur, g_prev = st[:, 0].copy(), st[:, 1].copy()
    w_cur, w_prev = st[:, 0].copy(), st[:, 1].copy()
    seen = set()
    for f in range(frames):
        bits = rng.integers(0, 2, size=(n, pb), dtype=np.uint8)
        if codec != 0:
            bits[: n // 2, 0] = 0          # keep half of the AMBE frames away from the tone / erasure ranges of b0
        status = dec.decode_parms(codec, bits, g_cur, g_prev)
        for i in range(n):
            d, c, p = vp(bits[i].ctypes.data), vp(w_cur[i].ctypes.data), vp(w_prev[i].ctypes.data)
            if codec == 0:
                rc = o.mbo_decode_imbe4400_parms(d, c, p)
            elif codec == 2:
                rc = o.mbo_decode_ambe2400_parms(d, c, p)
            else:
                rc = o.mbo_decode_ambe2450_parms(d, c, p, -1)
            assert status[i] == rc, (f, i)
            seen.add(int(rc))
        assert np.array_equal(g_cur, w_cur), "cur_mp differs after frame %d" % f
        assert np.array_equal(g_prev, w_prev), "prev_mp differs after frame %d" % f
        # mbe_moveMbeParms(cur, prev) for the streams that decoded a voice frame, like the frame paths do
        ok = status == 0
        g_prev[ok] = g_cur[ok]
        w_prev[ok] = w_cur[ok]
    assert 0 in seen and len(seen) >= 2        # voice frames and at least one special return code were exercised
    bad = rng.integers(0, 2, size=(2, pb), dtype=np.uint8)
    bad[1, 5] = 3
    before = g_cur[:2].copy()
    status = dec.decode_parms(codec, bad, g_cur[:2], g_prev[:2])
    assert status[1] == -2 and np.array_equal(g_cur[1], before[1])


@pytest.mark.parametrize("codec", [0, 3])
def test_enhance_and_smoothing(dec, codec):
    o = T.load_oracle()
    o.mbo_spectral_amp_enhance.restype = ctypes.c_float
    o.mbo_spectral_amp_enhance.argtypes = [vp]
    o.mbo_adaptive_smoothing.argtypes = [vp, vp, ctypes.c_int, ctypes.c_float]
    n = 128
    pb = T.PARAM_BITS[codec]
    rng = np.random.default_rng(0x9B0 + codec)
    st = _fresh(n)
    cur, prev, enh = st[:, 0].copy(), st[:, 1].copy(), st[:, 2].copy()
    bits = rng.integers(0, 2, size=(n, pb), dtype=np.uint8)
    bits[:, 0] = 0
    status = dec.decode_parms(codec, bits, cur, prev)
    keep = status == 0
    cur, enh = np.ascontiguousarray(cur[keep]), np.ascontiguousarray(enh[keep])
    m = cur.shape[0]
    assert m > n // 2
    # error statistics that switch the smoothing on for part of the batch (mbe_adaptive.c:151-276)
    for i in range(m):
        cur[i].view(np.float32)[293] = [0.0, 0.004, 0.01, 0.02, 0.05][i % 5]     # errorRate
        cur[i].view(np.int32)[294] = [0, 3, 5, 7, 9][(i // 5) % 5]                # errorCountTotal
        cur[i].view(np.int32)[295] = (i // 25) % 2                                # errorCount4
    want = cur.copy()
    rm0 = dec.spectral_amp_enhance(cur)
    for i in range(m):
        r = o.mbo_spectral_amp_enhance(vp(want[i].ctypes.data))
        assert np.float32(r).view(np.uint32) == rm0[i].view(np.uint32), i
    assert np.array_equal(cur, want), "mbe_spectralAmpEnhance"
    dec.adaptive_smoothing(cur, enh)
    for i in range(m):
        o.mbo_adaptive_smoothing(vp(want[i].ctypes.data), vp(enh[i].ctypes.data), 0, 0.0)
    assert np.array_equal(cur, want), "mbe_applyAdaptiveSmoothing"
