// Parameter dequantisation on one warp: parameter bits -> w0, L, K, Vl[], log2Ml[], Ml[] using the
// previous frame (replaces src/imbe/imbe7200x4400.c:117-354,589-630, src/ambe/ambe3600x2450.c:176-621,
// src/ambe/ambe3600x2400.c:164-546).  Lanes run over harmonics; sums whose order matters for
// rounding are kept in index order (per-harmonic terms are computed in parallel, then added serially).
#pragma once
#include "mbe_common.cuh"
#include "mbe_frontend.cuh"
#include "mbe_libm.cuh"

namespace mbe {

// codec tables (device copies of mbe_tables.inc) are declared in mbe_b200.cu before this header
__device__ __forceinline__ float pow2i(int e) {  // exact 2^e for -126 <= e <= 127 (exp2f of an integer)
    return __int_as_float((127 + e) << 23);
}

// Runs of consecutive parameter bits, MSB first, from the bit-reversed parameter words: frame bit i sits at bit
// 63 - i of (hi:lo) = (brev(dw[0]) : brev(dw[1])), so bits a .. a+n-1 are one funnel shift and a mask.
struct RevBits {
    unsigned hi, lo;
};
__device__ __forceinline__ RevBits rev_bits(const unsigned dw[3]) {
    RevBits r = {__brev(dw[0]), __brev(dw[1])};
    return r;
}
template <int A, int N>
__device__ __forceinline__ unsigned field(const RevBits& r) {
    static_assert(A >= 0 && N >= 1 && N <= 16 && A + N <= 64, "field range");
    constexpr int sh = 64 - A - N;
    const unsigned v = (sh >= 32) ? (r.hi >> (sh - 32)) : __funnelshift_r(r.lo, r.hi, sh);
    return v & ((1u << N) - 1u);
}

// log-magnitude prediction + exp2 (imbe7200x4400.c:294-354 / ambe3600x2450.c:389-459).
//   ambe = 0: rho = rho_imbe, clamp interpolation indices to 0..56, no gain term
//   ambe = 1: rho = 0.65, add BigGamma, unvoiced magnitudes scaled by unvc
__device__ __forceinline__ void predict_magnitudes(WarpWS& ws, const DevTables* T, int ambe, float rho, float unvc,
                                                   int lane) {
    ParmsSmall& cur = ws.cur;
    PrevSmall& prev = ws.prev;
    const int cur_L = cur.L;  // already within 9..56
    int prev_L = prev.L;
    prev_L = prev_L < 1 ? 1 : (prev_L > 56 ? 56 : prev_L);
    if (cur_L > prev_L) {
        const float m = prev.Ml[prev_L], lg = prev.log2Ml[prev_L];
        for (int l = prev_L + 1 + lane; l <= cur_L; l += 32) {
            prev.Ml[l] = m;
            prev.log2Ml[l] = lg;
        }
    }
    __syncwarp();
    if (lane == 0) {
        prev.log2Ml[0] = prev.log2Ml[1];
        prev.Ml[0] = prev.Ml[1];
    }
    __syncwarp();
    const float* P = prev.log2Ml;  // P[57] aliases PHIl[0], exactly as in the reference's struct
    const float ratio = (float)prev_L / (float)cur_L;
    // ordered sums over the harmonics: the terms of harmonic l sit at [l - 1], zero padded to a multiple of four harmonics
    // (x + 0 = x), so the serial loops move four terms per LDS.128.  AMBE: (interpolated term, Tl) pairs; the DCT input is dead.
    const int Lpad = (cur_L + 3) & ~3;
    float2* pair = reinterpret_cast<float2*>(ws.u.dec.tmp);
    float dl[2], pa[2], pb[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int l = 1 + lane + 32 * r;
        dl[r] = 0.f;
        pa[r] = 0.f;
        pb[r] = 0.f;
        if (l <= cur_L) {
            float fk = ratio * (float)l;
            int ik = (int)fk;
            int up;
            if (!ambe) {
                ik = ik < 0 ? 0 : (ik > 56 ? 56 : ik);
                up = ik + 1 > 56 ? 56 : ik + 1;
            } else {
                up = ik + 1;
            }
            dl[r] = fk - (float)ik;
            pa[r] = P[ik];
            pb[r] = P[up];
            const float term = (((float)1 - dl[r]) * pa[r]) + (dl[r] * pb[r]);
            if (ambe) {
                pair[l - 1] = make_float2(term, ws.u.dec.Tl[l]);
            } else {
                ws.u.dec.tmp[l - 1] = term;
            }
        } else if (l <= Lpad) {
            if (ambe) {
                pair[l - 1] = make_float2(0.f, 0.f);
            } else {
                ws.u.dec.tmp[l - 1] = 0.f;
            }
        }
    }
    __syncwarp();
    float acc = 0.f, s42 = 0.f;
    {
        const float4* p4 = reinterpret_cast<const float4*>(ws.u.dec.tmp);
        if (ambe) {
#pragma unroll 2
            for (int i = 0; i < Lpad; i += 4) {
                const float4 u = p4[i >> 1], v = p4[(i >> 1) + 1];
                acc = acc + u.x;
                s42 += u.y;
                acc = acc + u.z;
                s42 += u.w;
                acc = acc + v.x;
                s42 += v.y;
                acc = acc + v.z;
                s42 += v.w;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < Lpad; i += 4) {
                const float4 u = p4[i >> 2];
                acc = acc + u.x;
                acc = acc + u.y;
                acc = acc + u.z;
                acc = acc + u.w;
            }
        }
    }
    acc = ((rho / (float)cur_L) * acc);
    float big_gamma = 0.f;
    if (ambe) {
        s42 = s42 / (float)cur_L;
        big_gamma = cur.gamma - (0.5f * T->log2_int[cur_L]) - s42;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int l = 1 + lane + 32 * r;
        if (l <= cur_L) {
            float c1 = (rho * ((float)1 - dl[r]) * pa[r]);
            float c2 = (rho * dl[r] * pb[r]);
            float lg;
            if (ambe) {
                lg = ws.u.dec.Tl[l] + c1 + c2 - acc + big_gamma;
            } else {
                lg = ws.u.dec.Tl[l] + c1 + c2 - acc;
            }
            cur.log2Ml[l] = lg;
            float m = mbelibm::exp2f_glibc(lg, d_exp2_tab);
            if (ambe && cur.Vl[l] != 1) {
                m = unvc * m;
            }
            cur.Ml[l] = m;
        }
    }
    __syncwarp();
}

// per-block inverse DCT: ws.u.dec.tmp[64 + l] holds the DCT coefficients flattened by harmonic index,
// blocklen[0..nblk-1] the block sizes.  Writes ws.u.dec.Tl[1..L].
__device__ __forceinline__ void block_idct(WarpWS& ws, const DevTables* T, const int* blocklen, int nblk, int L, int lane) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int l = 1 + lane + 32 * r;
        if (l <= L) {
            int start = 0, ji = 1;
            for (int i = 0; i < nblk; ++i) {
                ji = blocklen[i];
                if (l <= start + ji) {
                    break;
                }
                start += ji;
            }
            const int j = l - start;
            const float* cs = T->blk + T->blk_off[ji] + (j - 1) * ji;
            float sum = 0.f;
            for (int k = 1; k <= ji; ++k) {
                float ak = (k == 1) ? 1.f : 2.f;
                sum = sum + (ak * ws.u.dec.tmp[64 + start + k] * cs[k - 1]);
            }
            ws.u.dec.Tl[l] = sum;
        }
    }
    __syncwarp();
}

// ---- IMBE 4400 ----  returns 0 (voice) or 1 (invalid fundamental -> repeat)
__device__ __forceinline__ int decode_imbe(const unsigned dw[3], WarpWS& ws, const DevTables* T, int lane) {
    ParmsSmall& cur = ws.cur;
    unsigned b0 = 0;
    {
        const int idx[8] = {0, 1, 2, 3, 4, 5, 85, 86};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            b0 = (b0 << 1) | getbit(dw, idx[i]);
        }
    }
    if (b0 > 207u) {
        return 1;
    }
    const int L = (int)T->imbe_L[b0];
    const int K = (int)T->imbe_Kv[b0];
    if (lane == 0) {
        cur.w0 = T->imbe_w0[b0];
        ws.w0row = (short)(COSW_IMBE + (int)b0);
        cur.L = L;
        cur.K = K;
    }
    const int L9 = L - 9;

    // scatter bits 6..84 into the quantiser words b1..bL+1
    for (int i = lane; i < 58; i += 32) {
        ws.u.dec.field[i] = 0;
    }
    __syncwarp();
    {
        const unsigned char* map = t_imbe_bitmap + L9 * 158;
        for (int i = 6 + lane; i < 85; i += 32) {
            if (getbit(dw, i)) {
                atomicOr(&ws.u.dec.field[map[2 * (i - 6)]], 1 << map[2 * (i - 6) + 1]);
            }
        }
    }
    __syncwarp();

    // voiced/unvoiced decisions: one bit per band of three harmonics
    {
        const int vbits = ws.u.dec.field[1];
        for (int l = 1 + lane; l <= L; l += 32) {
            int k = K - 1 - (l - 1) / 3;
            k = k < 0 ? 0 : k;
            cur.Vl[l] = (vbits >> k) & 1;
        }
    }
    // gain vector -> ws.u.dec.tmp[1..6]
    if (lane < 6) {
        float g;
        if (lane == 0) {
            g = t_imbe_gain0[ws.u.dec.field[2] & 63];
        } else {
            const int nb = t_imbe_gain_bits[L9 * 5 + (lane - 1)];
            const float step = t_imbe_gain_step[L9 * 5 + (lane - 1)];
            const int bm = ws.u.dec.field[lane + 2] & ((1 << nb) - 1);
            g = (step * ((float)bm - pow2i(nb - 1) + 0.5f));
        }
        ws.u.dec.tmp[1 + lane] = g;
    }
    __syncwarp();
    // 6-point inverse DCT of the gains -> ws.u.dec.tmp[9..14] = Ri[1..6]
    if (lane < 6) {
        float sum = 0.f;
#pragma unroll
        for (int m = 1; m <= 6; ++m) {
            float am = (m == 1) ? 1.f : 2.f;
            sum = sum + (am * ws.u.dec.tmp[m] * T->ri6[(m - 1) * 6 + lane]);
        }
        ws.u.dec.tmp[9 + lane] = sum;
    }
    __syncwarp();
    // DCT coefficients flattened by harmonic: first of each block = Ri, rest = dequantised HOCs
    int blen[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        blen[i] = t_imbe_blocklen[L9 * 6 + i];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int l = 1 + lane + 32 * r;
        if (l <= L) {
            int start = 0, blk = 0;  // block lengths are non-decreasing and sum to L
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (l > start + blen[i]) {
                    start += blen[i];
                    blk = i + 1;
                }
            }
            blk = blk > 5 ? 5 : blk;
            const int k = l - start;       // 1-based position inside block (blk+1)
            float v;
            if (k == 1) {
                v = ws.u.dec.tmp[9 + blk];
            } else {
                const int m = l - (blk + 1) + 7;  // quantiser word index
                const int Bm = t_imbe_hoc_bits[L9 * 50 + (m - 8)];
                if (Bm <= 0) {
                    v = 0.f;
                } else {
                    const int bm = ws.u.dec.field[m] & ((1 << Bm) - 1);
                    v = ((t_imbe_hoc_step[Bm - 1] * t_imbe_hoc_sdev[k - 2]) * (((float)bm - pow2i(Bm - 1)) + 0.5f));
                }
            }
            ws.u.dec.tmp[64 + l] = v;
        }
    }
    __syncwarp();
    block_idct(ws, T, blen, 6, L, lane);

    float rho;
    if (L <= 15) {
        rho = 0.4f;
    } else if (L <= 24) {
        rho = (0.03f * (float)L) - 0.05f;
    } else {
        rho = 0.7f;
    }
    predict_magnitudes(ws, T, 0, rho, 0.f, lane);
    return 0;
}

// ---- AMBE common tail: PRBA -> Ri -> Cik -> Tl -> magnitudes ----
struct AmbeBooks {
    const float* prba24;
    const float* prba58;
    const float* hoc[4];
    const unsigned char* blocklen;
};

__device__ __forceinline__ void ambe_tail(WarpWS& ws, const DevTables* T, const AmbeBooks& bk, int b3, int b4,
                                          const int hocidx[4], float unvc, int lane) {
    ParmsSmall& cur = ws.cur;
    const int L = cur.L;
    if (lane < 8) {
        float g;
        if (lane == 0) {
            g = 0.f;
        } else if (lane < 4) {
            g = bk.prba24[b3 * 3 + (lane - 1)];
        } else {
            g = bk.prba58[b4 * 4 + (lane - 4)];
        }
        ws.u.dec.tmp[1 + lane] = g;
    }
    __syncwarp();
    if (lane < 8) {
        float sum = 0.f;
#pragma unroll
        for (int m = 1; m <= 8; ++m) {
            float am = (m == 1) ? 1.f : 2.f;
            sum = sum + (am * ws.u.dec.tmp[m] * T->ri8[(m - 1) * 8 + lane]);
        }
        ws.u.dec.tmp[11 + lane] = sum;  // Ri[1..8] at tmp[11..18]
    }
    __syncwarp();
    int blen[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        blen[i] = bk.blocklen[L * 4 + i];
    }
    const float rconst = T->ambe_rconst;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int l = 1 + lane + 32 * r;
        if (l <= L) {
            int start = 0, blk = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (l > start + blen[i]) {
                    start += blen[i];
                    blk = i + 1;
                }
            }
            blk = blk > 3 ? 3 : blk;
            const int k = l - start;
            const float ra = ws.u.dec.tmp[11 + 2 * blk], rb2 = ws.u.dec.tmp[12 + 2 * blk];
            float v;
            if (k == 1) {
                v = 0.5f * (ra + rb2);
            } else if (k == 2) {
                v = rconst * (ra - rb2);
            } else if (k <= 6) {
                const float* h = (blk == 0) ? bk.hoc[0] : (blk == 1 ? bk.hoc[1] : (blk == 2 ? bk.hoc[2] : bk.hoc[3]));
                const int hi = (blk == 0) ? hocidx[0] : (blk == 1 ? hocidx[1] : (blk == 2 ? hocidx[2] : hocidx[3]));
                v = h[hi * 4 + (k - 3)];
            } else {
                v = 0.f;
            }
            ws.u.dec.tmp[64 + l] = v;
        }
    }
    __syncwarp();
    block_idct(ws, T, blen, 4, L, lane);
    predict_magnitudes(ws, T, 1, 0.65f, unvc, lane);
}

// ---- AMBE+2 3600x2450 ---- returns 0 voice | 2 erasure | 7 tone
__device__ __forceinline__ int decode_ambe2450(const unsigned dw[3], WarpWS& ws, const DevTables* T, int total_errors,
                                               int lane) {
    ParmsSmall& cur = ws.cur;
    // tone frame signature (ambe3600x2450.c:474-519) straight from the packed parameter words (bit i of the
    // frame = bit i & 31 of dw[i >> 5]): u0[11:6] = bits 0..5 all ones, u3[3:0] = bits 45..48 all zero, or the
    // two nibbles u1[11:8] = bits 12..15 and u1[3:0] = bits 20..23 equal
    const bool tone_ok = ((dw[0] & 0x3fu) == 0x3fu) &&
                         ((((dw[1] >> 13) & 0xfu) == 0u) || (((dw[0] >> 12) & 0xfu) == ((dw[0] >> 20) & 0xfu)));
    if (tone_ok && total_errors < 6) {
        return 7;
    }
    const RevBits rb = rev_bits(dw);
    const int b0 = (int)((field<0, 4>(rb) << 3) | field<37, 3>(rb));  // bits 0-3, 37-39
    if (b0 >= 120 && b0 <= 123) {
        return 2;
    }
    if (b0 == 126 || b0 == 127) {
        return 2;
    }
    int L, silence = 0;
    float f0, w0;
    if (b0 == 124 || b0 == 125) {
        silence = 1;
        f0 = T->a2450_f0_silence;
        w0 = T->a2450_w0_silence;
        L = (b0 == 124) ? 15 : 14;
    } else {
        f0 = t_a2450_f0[b0];
        w0 = T->a2450_w0[b0];
        L = t_a2450_L[b0];
    }
    const float unvc = 0.2046f / sqrtf(w0);
    const unsigned vmask = t_a2450_vuv[(field<4, 4>(rb) << 1) | field<35, 1>(rb)];   // bits 4-7, 35
    const float dg = t_a2450_dgain[(field<8, 4>(rb) << 1) | field<36, 1>(rb)];       // bits 8-11, 36
    for (int l = 1 + lane; l <= L; l += 32) {
        if (silence) {
            cur.Vl[l] = 0;
        } else {
            int jl = (int)((float)l * 16.0f * f0);
            cur.Vl[l] = (int)((vmask >> jl) & 1u);
        }
    }
    if (lane == 0) {
        cur.w0 = w0;
        ws.w0row = (short)(silence ? COSW_A2450_SILENCE : COSW_A2450 + b0);
        cur.L = L;
        cur.gamma = dg + (0.5f * ws.prev.gamma);
    }
    __syncwarp();
    AmbeBooks bk = {t_a2450_prba24, t_a2450_prba58, {t_a2450_hoc5, t_a2450_hoc6, t_a2450_hoc7, t_a2450_hoc8},
                    t_a2450_blocklen};
    // b5 = bits 24-27, 44; b6 = 28-30, 45; b7 = 31-33, 46; b8 = 34, 47, 48; b3 = 12-19, 40; b4 = 20-23, 41-43
    const int hocidx[4] = {(int)((field<24, 4>(rb) << 1) | field<44, 1>(rb)), (int)((field<28, 3>(rb) << 1) | field<45, 1>(rb)),
                           (int)((field<31, 3>(rb) << 1) | field<46, 1>(rb)), (int)((field<34, 1>(rb) << 2) | field<47, 2>(rb))};
    ambe_tail(ws, T, bk, (int)((field<12, 8>(rb) << 1) | field<40, 1>(rb)), (int)((field<20, 4>(rb) << 3) | field<41, 3>(rb)),
              hocidx, unvc, lane);
    return 0;
}

// ---- AMBE 3600x2400 ---- returns 0 voice | 3 tone/silence marker | 5..122 D-STAR tone index
__device__ __forceinline__ int decode_ambe2400(const unsigned dw[3], WarpWS& ws, const DevTables* T, int lane) {
    ParmsSmall& cur = ws.cur;
    const RevBits rb = rev_bits(dw);
    const int b0 = (int)((field<0, 6>(rb) << 1) | field<48, 1>(rb));  // bits 0-5, 48
    if ((b0 & 0x7E) == 0x7E) {
        // three remapped high bits (t7,t6,t5 tables of the reference folded into one) + five literal bits
        const unsigned hi3 = (0x56732104u >> (4 * ((getbit(dw, 6) << 2) | (getbit(dw, 7) << 1) | getbit(dw, 8)))) & 7u;
        const int tone = (int)((hi3 << 5) | (getbit(dw, 9) << 4) | (getbit(dw, 42) << 3) | (getbit(dw, 43) << 2)
                               | (getbit(dw, 10) << 1) | getbit(dw, 11));
        if (tone >= 5 && tone <= 122) {
            return tone;
        }
        if (!(tone >= 128 && tone <= 163)) {
            for (int l = 1 + lane; l <= 14; l += 32) {
                cur.Vl[l] = 0;
            }
            if (lane == 0) {
                cur.w0 = T->a2400_w0_silence;
                cur.L = 14;
            }
            __syncwarp();
        }
        return 3;
    }
    const float f0 = T->a2400_f0[b0];
    const float w0 = T->a2400_w0[b0];
    const int L = t_a2400_L[b0];
    const float unvc = 0.2046f / sqrtf(w0);
    const unsigned vmask = t_a2400_vuv[field<38, 4>(rb)];                              // bits 38-41
    const float dg = t_a2400_dgain[(field<6, 4>(rb) << 2) | field<42, 2>(rb)];         // bits 6-9, 42, 43
    for (int l = 1 + lane; l <= L; l += 32) {
        int jl = (int)((float)l * 16.0f * f0);
        cur.Vl[l] = (int)((vmask >> jl) & 1u);
    }
    if (lane == 0) {
        cur.w0 = w0;
        ws.w0row = (short)(COSW_A2400 + b0);
        cur.L = L;
        cur.gamma = dg + (0.5f * ws.prev.gamma);
    }
    __syncwarp();
    AmbeBooks bk = {t_a2400_prba24, t_a2400_prba58, {t_a2400_hoc5, t_a2400_hoc6, t_a2400_hoc7, t_a2400_hoc8},
                    t_a2400_blocklen};
    // b5 = bits 22, 23, 25, 26; b6 = 27-30; b7 = 31-34; b8 = 35-37 (LSB forced 0); b3 = 10-16, 44, 45; b4 = 17-21, 46, 47
    const int hocidx[4] = {(int)((field<22, 2>(rb) << 2) | field<25, 2>(rb)), (int)field<27, 4>(rb), (int)field<31, 4>(rb),
                           (int)(field<35, 3>(rb) << 1)};
    ambe_tail(ws, T, bk, (int)((field<10, 7>(rb) << 2) | field<44, 2>(rb)), (int)((field<17, 5>(rb) << 2) | field<46, 2>(rb)),
              hocidx, unvc, lane);
    return 0;
}

}  // namespace mbe
