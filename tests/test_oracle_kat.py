"""Known-answer tests that pin the CPU oracle (oracle/mbe_oracle.c) to the vectors the reference's OWN tests
hold for this path, independent of oracle/_ref being present:

  tests/test_golden_pcm.c:49-84      one synthetic frame -> FNV-1a float 0x59741032, int16 0x4EDB8636
  tests/test_ecc.c:221-406           Hamming(15,11) both layouts, Golay(23,12) hard + soft vectors
  tests/test_floattoshort_parity.c   clip edges, NaN -> 0, +-Inf -> +-clip, truncation
  tests/test_noise_determinism.c     identical state => identical output; noiseSeed advances
"""
import ctypes

import numpy as np
import pytest

import mbe_testlib as T

GOLAY_GEN = T._GOLAY_GEN
HAM_STD = [0x7f08, 0x78e4, 0x66d2, 0x55b1]
HAM_7100 = [0x7ac8, 0x3d64, 0x1eb2, 0x7591]


def _o():
    return T.load_oracle()


def _golden_params():
    o = _o()
    cur = np.zeros(T.PARMS_BYTES, np.uint8)
    prev = np.zeros(T.PARMS_BYTES, np.uint8)
    enh = np.zeros(T.PARMS_BYTES, np.uint8)
    o.mbo_init_parms(T._ptr(cur), T._ptr(prev), T._ptr(enh))
    f, i = cur.view(np.float32), cur.view(np.int32)
    f[0] = np.float32(0.105)
    i[1] = 36
    for l in range(1, 37):
        i[3 + l] = 1 if l % 4 else 0
        f[60 + l] = np.float32(0.035) + np.float32(0.0015) * np.float32(l)
        f[174 + l] = np.float32(l) * np.float32(0.03)
        f[231 + l] = np.float32(l) * np.float32(0.02)
    prev[:] = cur
    return cur, prev


def test_golden_pcm_hashes():
    """tests/test_golden_pcm.c: fill_params (:49-62), seed 0xC0FFEE, mbe_synthesizeSpeechf + mbe_floattoshort."""
    o = _o()
    cur, prev = _golden_params()
    rng = np.zeros(16, np.uint8)
    o.mbo_rng_default(T._ptr(rng))
    o.mbo_rng_seed(T._ptr(rng), ctypes.c_uint32(0xC0FFEE))
    out = np.zeros(160, np.float32)
    pcm = np.zeros(160, np.int16)
    o.mbo_synthesize_speech(T._ptr(out), T._ptr(cur), T._ptr(prev), 0, ctypes.c_float(0.0), T._ptr(rng))
    o.mbo_float_to_short(T._ptr(out), T._ptr(pcm))
    assert T.fnv1a32(out.tobytes()) == 0x59741032
    assert T.fnv1a32(pcm.tobytes()) == 0x4EDB8636


def _bits(word, n):
    return np.array([(word >> j) & 1 for j in range(n)], np.int8)


def _soft(bits, weak):
    s = np.zeros((len(bits), 2), np.uint8)
    s[:, 0] = bits
    s[:, 1] = 200
    for w in weak:
        s[w, 1] = 1
    return s


def _ham_encode(rows, data_pos, data11):
    code = 0
    for i, p in enumerate(data_pos):
        code |= int(data11[i]) << p
    parity_pos = [p for p in range(15) if p not in data_pos]
    for par in range(16):
        c = code
        for k, p in enumerate(parity_pos):
            c |= ((par >> k) & 1) << p
        if all(bin(c & r).count("1") % 2 == 0 for r in rows):
            return c
    raise AssertionError


def _hamming_case(variant, rows, data_pos, data11):
    o = _o()
    cw = _ham_encode(rows, data_pos, data11)
    code = _bits(cw, 15)
    out = np.zeros(15, np.int8)
    assert o.mbo_hamming1511(T._ptr(code), T._ptr(out), variant) == 0      # fixed point
    assert np.array_equal(out, code)
    for k in range(15):                                                      # every single-bit flip is corrected
        err = code.copy()
        err[14 - k] ^= 1
        assert o.mbo_hamming1511(T._ptr(err), T._ptr(out), variant) >= 1
        assert np.array_equal(out[data_pos], code[data_pos])
        if variant == 1:
            assert np.array_equal(out, code)
    err = code.copy()                                                        # two weak flips: soft decode returns 2
    err[2] ^= 1
    err[4] ^= 1
    soft = _soft(err, [2, 4])
    assert o.mbo_hamming1511_soft(T._ptr(soft), T._ptr(out), variant) == 2
    assert np.array_equal(out, code)


def test_hamming_standard_vectors():
    """tests/test_ecc.c:221-272 (data pattern i % 2, data positions of the standard layout)."""
    _hamming_case(0, HAM_STD, [2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14], [(i % 2) == 0 for i in range(11)])


def test_hamming_7100_vectors():
    """tests/test_ecc.c:274-352 incl. all 2048 clean codewords soft-decoding clean."""
    pos = [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    _hamming_case(1, HAM_7100, pos, [(i % 3) == 0 for i in range(11)])
    o = _o()
    out = np.zeros(15, np.int8)
    for data in range(2048):
        cw = _ham_encode(HAM_7100, pos, [(data >> i) & 1 for i in range(11)])
        clean = _bits(cw, 15)
        assert o.mbo_hamming1511_soft(T._ptr(_soft(clean, [])), T._ptr(out), 1) == 0
        assert np.array_equal(out, clean)


def test_golay_vectors():
    """tests/test_ecc.c:354-406: single error via mbe_checkGolayBlock; soft 4 weak flips -> 2; parity echo."""
    o = _o()
    data = 0xA55
    ecc = 0
    for i in range(12):
        if (data >> (11 - i)) & 1:
            ecc ^= GOLAY_GEN[i]
    block = (data << 11) | ecc
    b = ctypes.c_long(block ^ (1 << 5))
    assert o.mbo_check_golay_block(ctypes.byref(b)) == 0
    assert b.value == data
    code = _bits(block, 23)
    out = np.zeros(23, np.int8)
    err = code.copy()
    for j in (22, 17, 8, 2):
        err[j] ^= 1
    assert o.mbo_golay2312_soft(T._ptr(_soft(err, [22, 17, 8, 2])), T._ptr(out)) == 2
    assert np.array_equal(out[11:], code[11:])
    err = code.copy()
    err[5] ^= 1
    assert o.mbo_golay2312_soft(T._ptr(_soft(err, [5])), T._ptr(out)) == 0
    assert np.array_equal(out[11:], code[11:])
    assert np.array_equal(out[:11], err[:11])          # parity bits echo the input
    # hard decoder: every single-bit error of every 64th data word is corrected
    for d in range(0, 4096, 64):
        cw = T.golay_encode(d)
        for j in range(23):
            e = _bits(cw ^ (1 << j), 23)
            n = o.mbo_golay2312(T._ptr(e), T._ptr(out))
            assert np.array_equal(out[11:], _bits(cw, 23)[11:])
            assert n == (1 if j >= 11 else 0)


def test_floattoshort_edges():
    """tests/test_floattoshort_parity.c:20-59 over its four LCG-seeded buffers."""
    o = _o()
    clip = np.float32((32767.0 * 0.95) / 7.0)
    for seed in (1, 0x12345678, 0xDEADBEEF, 0xC0FFEE):
        x = np.zeros(160, np.float32)
        st = seed
        for i in range(160):
            st = (st * 1664525 + 1013904223) & 0xffffffff
            x[i] = np.float32(((st >> 8) - 0x007FFFFF)) / np.float32(65536.0)
        d = np.float32(1.0 / 32768.0)
        x[:12] = [0.0, clip, clip + d, clip - d, -clip, -clip - d, -clip + d, np.float32(1.0) / np.float32(7.0),
                  np.float32(-1.0) / np.float32(7.0), np.nan, np.inf, -np.inf]
        got = np.zeros(160, np.int16)
        o.mbo_float_to_short(T._ptr(x), T._ptr(got))
        maxa = np.float32(32767.0) * np.float32(0.95)
        with np.errstate(invalid="ignore"):
            a = np.float32(7.0) * x
            a = np.where(np.isnan(a), np.float32(0), a)
            a = np.clip(a, -maxa, maxa)
        want = np.trunc(a).astype(np.int16)
        assert np.array_equal(got, want)


def test_noise_determinism():
    """tests/test_noise_determinism.c: all-unvoiced L = 24 frame."""
    o = _o()

    def fresh():
        cur = np.zeros(T.PARMS_BYTES, np.uint8)
        prev = cur.copy()
        enh = cur.copy()
        o.mbo_init_parms(T._ptr(cur), T._ptr(prev), T._ptr(enh))
        f, i = cur.view(np.float32), cur.view(np.int32)
        f[0] = np.float32(0.10)
        i[1] = 24
        for l in range(1, 25):
            i[3 + l] = 0
            f[60 + l] = np.float32(0.04) + np.float32(0.001) * np.float32(l)
            f[174 + l] = 0
            f[231 + l] = 0
        prev[:] = cur
        rng = np.zeros(16, np.uint8)
        o.mbo_rng_default(T._ptr(rng))
        return cur, prev, rng

    def synth(cur, prev, rng):
        out = np.zeros(160, np.float32)
        o.mbo_synthesize_speech(T._ptr(out), T._ptr(cur), T._ptr(prev), 0, ctypes.c_float(0.0), T._ptr(rng))
        return out

    c1, p1, r1 = fresh()
    c2, p2, r2 = fresh()
    a, b = synth(c1, p1, r1), synth(c2, p2, r2)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    seed_after_first = T.parms_view(c1)["noiseSeed"]
    p1[:] = c1                                           # mbe_moveMbeParms(cur -> prev)
    c = synth(c1, p1, r1)
    assert not np.array_equal(a, c)
    assert T.parms_view(c1)["noiseSeed"] != seed_after_first


def test_fft256_roundtrip_and_dc():
    """the 256-point FFTPACK restatement: backward(forward(x)) == 256 x (unnormalised pair), DC bin = sum."""
    o = _o()
    rng = np.random.default_rng(1)
    x = rng.normal(0, 1, 256).astype(np.float32)
    X = np.zeros(256, np.float32)
    y = np.zeros(256, np.float32)
    o.mbo_fft256_forward_ordered(T._ptr(x), T._ptr(X))
    o.mbo_fft256_backward_ordered(T._ptr(X), T._ptr(y))
    assert np.allclose(y / 256.0, x, atol=2e-5)
    ref = np.fft.rfft(x.astype(np.float64))
    assert abs(X[0] - ref[0].real) < 1e-3 and abs(X[1] - ref[128].real) < 1e-3
    assert np.allclose(X[2::2], ref[1:128].real, atol=2e-3) and np.allclose(X[3::2], ref[1:128].imag, atol=2e-3)


@pytest.mark.parametrize("code", [0, 1, 2])
def test_fast_exhaustive_soft_decoders_equal_the_restatement(code):
    """oracle/mbe_oracle.c has two exhaustive soft decoders: the line-by-line restatement of src/ecc/ecc.c:157-215,303-357
    (pinned to the compiled reference by test_oracle_vs_ref.py and to tests/test_ecc.c by the vectors above) and a table
    driven one (same enumeration order, same tie-break function) that the million-word GPU differential uses.  They must
    agree word for word, including on reliabilities that force cost ties."""
    words = T.ecc_test_words(code, 6000, 0xECC0 + code)
    a = T.oracle_ecc_blocks(code, words, soft=True, fast=False)
    b = T.oracle_ecc_blocks(code, words, soft=True, fast=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (a[1] > 0).mean() > 0.3      # the sample does exercise corrections
