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


STEP_FN = {0: ("mbe_eccImbe7200x4400C0", "mbe_demodulateImbe7200x4400Data", "mbe_eccImbe7200x4400Data"),
           1: ("mbe_eccImbe7100x4400C0", "mbe_demodulateImbe7100x4400Data", "mbe_eccImbe7100x4400Data"),
           2: ("mbe_eccAmbe3600x2400C0", "mbe_demodulateAmbe3600x2400Data", "mbe_eccAmbe3600x2400Data"),
           3: ("mbe_eccAmbe3600x2450C0", "mbe_demodulateAmbe3600x2450Data", "mbe_eccAmbe3600x2450Data")}


@pytest.mark.parametrize("codec", [0, 1, 2, 3])
def test_channel_steps(dec, codec):
    """SURVEY 8(a) rows E6-E9 step by step: C0 ECC, de-scrambling, data ECC (and the 7100 -> 7200 bit permutation) run
    one at a time on caller-held frames.  The chain must equal the fused front-end (mbe_b200_decode_frames), and - when
    the compiled reference travels with the repo (oracle/_ref) - every intermediate frame must equal what the reference's
    own mbe_ecc<Codec>C0 / mbe_demodulate<Codec>Data / mbe_ecc<Codec>Data / mbe_convertImbe7100to7200 leave behind."""
    n = 200
    fb, pb = T.FRAME_BITS[codec], T.PARAM_BITS[codec]
    frames = T.random_hard_frames(codec, n, 1, 0x5E0 + codec).reshape(n, fb)
    want_bits, want_res = dec.decode_frames(codec, frames)
    ref = T.load_ref_api()
    g = frames.copy()
    r = frames.copy()
    c0 = dec.channel_step(codec, 0, frames=g)
    if ref is not None:
        rc0 = np.array([getattr(ref, STEP_FN[codec][0])(vp(r[i].ctypes.data)) for i in range(n)])
        assert np.array_equal(c0, rc0) and np.array_equal(g, r), "C0 step"
    st = dec.channel_step(codec, 1, frames=g)
    assert not st.any()
    if ref is not None:
        for i in range(n):
            assert getattr(ref, STEP_FN[codec][1])(vp(r[i].ctypes.data)) == 0
        assert np.array_equal(g, r), "de-scramble step"
    bits = np.zeros((n, pb), np.uint8)
    before = g.copy()
    prot = dec.channel_step(codec, 2, frames=g, bits=bits)
    assert np.array_equal(g, before)                      # the data step does not modify the frame
    if ref is not None:
        rbits = np.zeros((n, pb), np.uint8)
        rprot = np.array([getattr(ref, STEP_FN[codec][2])(vp(r[i].ctypes.data), vp(rbits[i].ctypes.data)) for i in range(n)])
        assert np.array_equal(prot, rprot) and np.array_equal(bits, rbits), "data step"
    if codec == 1:
        st = dec.channel_step(codec, 3, bits=bits)
        assert not st.any()
        if ref is not None:
            for i in range(n):
                assert ref.mbe_convertImbe7100to7200(vp(rbits[i].ctypes.data)) == 0
            assert np.array_equal(bits, rbits), "7100 -> 7200 permutation"
    assert np.array_equal(bits, want_bits)
    assert np.array_equal(c0, want_res["c0_errors"]) and np.array_equal(prot, want_res["protected_errors"])
    bad = frames[:3].copy()
    bad[1, 9] = 7
    keep = bad.copy()
    st = dec.channel_step(codec, 0, frames=bad)
    assert st[1] == -2 and np.array_equal(bad[1], keep[1])


def test_tone_and_comfort_noise(dec):
    """Rows S5 / S6 on their own: mbe_synthesizeTonef, mbe_synthesizeTonefdstar and mbe_synthesizeComfortNoisef against
    the oracle, float samples bitwise, tone phases and RNG state carried from call to call."""
    o = T.load_oracle()
    o.mbo_synthesize_tone.argtypes = [vp, vp, vp]
    o.mbo_synthesize_tone_dstar.argtypes = [vp, vp, ctypes.c_int]
    o.mbo_comfort_noise.argtypes = [vp, vp]
    rng = np.random.default_rng(0x70E)
    ids = [5, 6, 7, 40, 122, 128, 141, 163, 4, 123, 200, 0]
    n = len(ids)
    bits = rng.integers(0, 2, size=(n, 49), dtype=np.uint8)
    for i, tid in enumerate(ids):             # ID1 = u1[11:4] = bits 12..19, AD from u0[5:0] (bits 6..11) and u3 bit 4
        for k in range(8):
            bits[i, 12 + k] = (tid >> (7 - k)) & 1
    st = _fresh(n)
    cur_g, cur_w = st[:, 0].copy(), st[:, 0].copy()
    out_w = np.zeros((n, 160), np.float32)
    for rep in range(3):                      # phases continue across frames
        out_g = dec.synthesize_tone(cur_g, bits49=bits)
        for i in range(n):
            o.mbo_synthesize_tone(vp(out_w[i].ctypes.data), vp(bits[i].ctypes.data), vp(cur_w[i].ctypes.data))
        assert np.array_equal(out_g.view(np.uint32), out_w.view(np.uint32)), "mbe_synthesizeTonef rep %d" % rep
        assert np.array_equal(cur_g, cur_w)
    assert out_g[:8].any() and not out_g[8:].any()        # valid ids sound, invalid ids are silent
    dst = np.array([5, 6, 64, 122, 128, 4, 123, -1], np.int32)
    m = len(dst)
    cur_g, cur_w = st[:m, 0].copy(), st[:m, 0].copy()
    for rep in range(2):
        out_g = dec.synthesize_tone(cur_g, dstar_id=dst)
        for i in range(m):
            o.mbo_synthesize_tone_dstar(vp(out_w[i].ctypes.data), vp(cur_w[i].ctypes.data), int(dst[i]))
        assert np.array_equal(out_g.view(np.uint32), out_w[:m].view(np.uint32)), "mbe_synthesizeTonefdstar"
        assert np.array_equal(cur_g, cur_w)
    # comfort noise: per-element RNG words (seeded like mbe_setThreadRngSeed), three frames in a row
    d2 = load_package().Decoder(max_streams=16, device=0)
    d2.init_streams(0, 16, np.arange(16, dtype=np.uint32) + 77)
    words_g = d2.export_rng(0, 16)
    d2.close()
    words_w = words_g.copy()
    noise_w = np.zeros((16, 160), np.float32)
    for rep in range(3):
        out_g = dec.comfort_noise(words_g)
        for i in range(16):
            o.mbo_comfort_noise(vp(noise_w[i].ctypes.data), vp(words_w[i].ctypes.data))
        assert np.array_equal(out_g.view(np.uint32), noise_w.view(np.uint32)), "mbe_synthesizeComfortNoisef"
        assert np.array_equal(words_g, words_w)
