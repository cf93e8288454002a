"""SURVEY 8(a) rows E1-E5 on their own: the block decoders of the CUDA front-end (mbe_b200_ecc_blocks = batched
mbe_golay2312[Soft], mbe_hamming1511[Soft], mbe_7100x4400hamming1511[Soft]) against the reference's known answers
(tests/test_ecc.c:221-406) and, word by word, against the oracle on random words and reliabilities."""
import ctypes

import numpy as np
import pytest

import mbe_testlib as T
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu

HAM = {1: ([0x7f08, 0x78e4, 0x66d2, 0x55b1], [2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14]),
       2: ([0x7ac8, 0x3d64, 0x1eb2, 0x7591], [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14])}


@pytest.fixture(scope="module")
def dec():
    d = load_package().Decoder(max_streams=1, device=0)
    yield d
    d.close()


def _bits(word, n):
    return np.array([(word >> i) & 1 for i in range(n)], np.uint8)


def _soft(bits, weak, strong=200):
    s = np.zeros(bits.shape + (2,), np.uint8)
    s[..., 0] = bits
    s[..., 1] = strong
    for w in weak:
        s[w, 1] = 1
    return s


def _ham_encode(code, data11):
    rows, pos = HAM[code]
    cw = 0
    for i, p in enumerate(pos):
        cw |= int(data11[i]) << p
    par = [p for p in range(15) if p not in pos]
    for v in range(16):
        c = cw
        for k, p in enumerate(par):
            c |= ((v >> k) & 1) << p
        if all(bin(c & r).count("1") % 2 == 0 for r in rows):
            return c
    raise AssertionError


@pytest.mark.parametrize("code", [1, 2])
def test_hamming_known_answers(dec, code):
    """tests/test_ecc.c:221-352: fixed point, every single flip corrected, two weak flips soft-decode to 2, and (7100
    layout) all 2048 clean codewords soft-decode clean."""
    pos = HAM[code][1]
    cw = _ham_encode(code, [(i % (2 if code == 1 else 3)) == 0 for i in range(11)])
    clean = _bits(cw, 15)
    words = np.stack([clean] + [clean ^ _bits(1 << k, 15) for k in range(15)])
    out, st = dec.ecc_blocks(code, words)
    assert st[0] == 0 and np.array_equal(out[0], clean)
    assert (st[1:] >= 1).all() and all(np.array_equal(o[pos], clean[pos]) for o in out[1:])
    err = clean.copy()
    err[2] ^= 1
    err[4] ^= 1
    out, st = dec.ecc_blocks(code, _soft(err, [2, 4])[None], soft=True)
    assert st[0] == 2 and np.array_equal(out[0], clean)
    allcw = np.stack([_bits(_ham_encode(code, [(d >> i) & 1 for i in range(11)]), 15) for d in range(2048)])
    out, st = dec.ecc_blocks(code, _soft(allcw, []), soft=True)
    assert (st == 0).all() and np.array_equal(out, allcw)


def test_golay_known_answers(dec):
    """tests/test_ecc.c:354-406: four weak flips soft-decode to 2 data-bit changes, parity bits echo the input, the hard
    decoder corrects every single-bit error of every 64th data word."""
    code = _bits(T.golay_encode(0xA55), 23)
    err = code.copy()
    for j in (22, 17, 8, 2):
        err[j] ^= 1
    out, st = dec.ecc_blocks(0, _soft(err, [22, 17, 8, 2])[None], soft=True)
    assert st[0] == 2 and np.array_equal(out[0][11:], code[11:])
    err = code.copy()
    err[5] ^= 1
    out, st = dec.ecc_blocks(0, _soft(err, [5])[None], soft=True)
    assert st[0] == 0 and np.array_equal(out[0][11:], code[11:]) and np.array_equal(out[0][:11], err[:11])
    words, want, nerr = [], [], []
    for d in range(0, 4096, 64):
        cw = T.golay_encode(d)
        for j in range(23):
            words.append(_bits(cw ^ (1 << j), 23))
            want.append(_bits(cw, 23)[11:])
            nerr.append(1 if j >= 11 else 0)
    out, st = dec.ecc_blocks(0, np.stack(words))
    assert np.array_equal(out[:, 11:], np.stack(want)) and np.array_equal(st, np.array(nerr))


@pytest.mark.parametrize("code", [0, 1, 2])
def test_random_words_equal_the_oracle(dec, code):
    o = T.load_oracle()
    vp = ctypes.c_void_p
    n, ln = 300, (23 if code == 0 else 15)
    rng = np.random.default_rng(40 + code)
    words = rng.integers(0, 2, size=(n, ln), dtype=np.uint8)
    levels = [rng.integers(0, 256, size=(n, ln)), rng.integers(0, 2, size=(n, ln)) * 255, rng.integers(0, 3, size=(n, ln))]
    out_h, st_h = dec.ecc_blocks(code, words)
    want = np.zeros(ln, np.int8)
    for i in range(n):
        if code == 0:
            rc = o.mbo_golay2312(vp(words[i].ctypes.data), vp(want.ctypes.data))
        else:
            rc = o.mbo_hamming1511(vp(words[i].ctypes.data), vp(want.ctypes.data), code - 1)
        assert rc == st_h[i] and np.array_equal(out_h[i], want.view(np.uint8)), i
    for rel in levels:
        soft = np.ascontiguousarray(np.stack([words, rel.astype(np.uint8)], axis=-1))
        out_s, st_s = dec.ecc_blocks(code, soft, soft=True)
        for i in range(n):
            if code == 0:
                rc = o.mbo_golay2312_soft(vp(soft[i].ctypes.data), vp(want.ctypes.data))
            else:
                rc = o.mbo_hamming1511_soft(vp(soft[i].ctypes.data), vp(want.ctypes.data), code - 1)
            assert rc == st_s[i] and np.array_equal(out_s[i], want.view(np.uint8)), i


@pytest.mark.parametrize("code", [0, 1, 2])
def test_soft_decoders_million_words_per_code(dec, code):
    """The soft decoders were restructured (coset walk, complement rule, lower-bound skip: mbe_frontend.cuh), which is where
    the reference's three-level tie-break (/root/reference/src/ecc/ecc.c:54-67,196-197) could silently break: 1,048,576 words
    per code - valid code words with 0..6 flips, five reliability families built to force ties (T.ecc_test_words) - through
    mbe_b200_ecc_blocks(soft=1) against the oracle's exhaustive search, every word; a 20,000-word subsample also against the
    line-by-line restatement."""
    n = 1 << 20
    words = T.ecc_test_words(code, n, 0x50F7ECC + code)
    got, st = dec.ecc_blocks(code, words, soft=True)
    want, wst = T.oracle_ecc_blocks(code, words, soft=True, fast=True, n_threads=16)
    bad = np.nonzero((st != wst) | (got != want).any(axis=1))[0]
    assert bad.size == 0, "first mismatch at word %d: %r" % (bad[0], words[bad[0]].tolist())
    sub = np.random.default_rng(code).choice(n, 20000, replace=False)
    want2, wst2 = T.oracle_ecc_blocks(code, words[sub], soft=True, fast=False, n_threads=16)
    assert np.array_equal(got[sub], want2) and np.array_equal(st[sub], wst2)
    hard, hst = dec.ecc_blocks(code, np.ascontiguousarray(words[..., 0]))
    print("code %d: %.1f %% of the soft decodes differ from the hard decode" % (code, 100.0 * (got != hard).any(axis=1).mean()))


def test_invalid_bits_leave_the_output_alone(dec):
    words = np.zeros((3, 23), np.uint8)
    words[1, 7] = 2
    out, st = dec.ecc_blocks(0, words)
    assert list(st) == [0, -2, 0] and not out.any()
