// Channel front-end on one warp: frame bits -> corrected, de-scrambled parameter bits.
// Replaces src/ecc/ecc.c, src/imbe/imbe7200x4400.c:424-778, src/imbe/imbe7100x4400.c:99-516 and
// src/ambe/ambe_common.c:22-189 of the reference.  Integer-only, bit-exact by construction.
//
// Layout: a row of the frame lives in ONE 32-bit word built with __ballot_sync (lane j contributes
// column j), so a Golay/Hamming decode is a handful of scalar integer instructions, the PN mask of a
// row is one ballot of per-lane jump-ahead LCG bits, and the soft-decision search is a warp-wide
// minimum over a packed (cost, tie-break) key.
#pragma once
#include "mbe_common.cuh"

namespace mbe {

struct FrontResult {
    int status;       // >= 0: total errors; < 0: MBE_STATUS_*
    int c0, prot, c4;
    unsigned flags;
};

__device__ __forceinline__ unsigned golay_parity(unsigned data12, const DevTables* T) {
    return (unsigned)T->golay_par_hi[data12 >> 6] ^ (unsigned)T->golay_par_lo[data12 & 63u];
}

// hard-decision Golay(23,12): returns corrected 12 data bits, *errs = number of changed data bits
// (src/ecc/ecc.c:221-301)
__device__ __forceinline__ unsigned golay_hard(unsigned w23, const DevTables* T, int* errs) {
    unsigned data = (w23 >> 11) & 0xfffu;
    unsigned syn = golay_parity(data, T) ^ (w23 & 0x7ffu);
    unsigned fixed = data ^ (unsigned)T->golay_fix[syn];
    *errs = __popc(fixed ^ data);
    return fixed;
}

__device__ __forceinline__ unsigned hamming_hard(unsigned w15, int variant, const DevTables* T, int* errs) {
    int syn = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        syn |= (__popc(w15 & (unsigned)T->ham_rows[variant][i]) & 1) << i;
    }
    *errs = syn > 0 ? 1 : 0;
    return w15 ^ (unsigned)T->ham_flip[variant][syn];
}

// Soft-decision decode of one row: exact maximum-likelihood search over all codewords with the reference's
// tie-break order (src/ecc/ecc.c:54-67,157-215): lowest cost, then the candidate equal to the hard decode, then
// fewest differing bits, then lowest data index.  Cost = sum of reliabilities of the positions where the candidate
// differs from the received hard bits.
//
// The search is exhaustive like the reference's, but it walks the coset instead of re-encoding every candidate:
// with d' = data ^ received data, the difference pattern of a candidate is (d', parity(d') ^ syndrome), so
//   key(d') = cost << 16 | differing bits << 12 | data  =  KA[d' >> 6] + KB[d' & 63] + CP[par(d' >> 6) ^ par(d' & 63) ^ syn] << 16
// is a sum of three table terms (every field is additive and cannot carry).  A lane keeps KB / par of its two low
// halves in registers and walks the high halves: one 16-bit shared-memory load and a few integer instructions per
// candidate; the complement of a candidate is a candidate as well (all-ones is a codeword) and its key is a constant
// minus the key, so only half of the high halves are walked, tracking minimum and maximum.  The hard decode is the only candidate that wins ties on "equal to the hard decode", so it is
// the answer unless some candidate costs strictly less than it does; and when the (dmin - t) least reliable
// positions outside its t corrected positions already weigh at least as much as those t, nothing can cost less
// and the search is skipped altogether.
struct SoftScratch {
    unsigned short* cp;  // 2048 x uint16 (4-byte aligned): cost of every 11-bit parity difference pattern
    unsigned* ka;        // 64 x uint32: partial key of each high data half
    unsigned short* qa;  // 64 x uint16: (parity of the high data half ^ syndrome) as a byte offset into cp
                         // (the first 32 / 16 entries of ka and qa are used: see the complement rule below)
};

// sum of the k least reliable positions outside `inside` (bit p = position p), nbits positions in all; stops early
// once the partial sum has reached `enough` (all the caller asks is whether the sum gets that far)
__device__ __forceinline__ unsigned soft_least_outside(const unsigned char* rel, unsigned inside, int nbits, int k,
                                                       unsigned enough, int lane) {
    unsigned mine = (lane < nbits && !((inside >> lane) & 1u)) ? (((unsigned)rel[lane] << 5) | (unsigned)lane) : 0xffffffffu;
    unsigned sum = 0;
    for (int i = 0; i < k && sum < enough; ++i) {
        const unsigned m = __reduce_min_sync(FULL, mine);
        sum += m >> 5;
        if (mine == m) {
            mine = 0xffffffffu;
        }
    }
    return sum;
}

__device__ __forceinline__ unsigned soft_cost_of(const unsigned char* rel, unsigned pattern, int nbits, int lane) {
    return __reduce_add_sync(FULL, (lane < nbits && ((pattern >> lane) & 1u)) ? (unsigned)rel[lane] : 0u);
}

// The search proper: the smallest key over the whole coset (see above).  The soft decoders are one copy per kernel
// (golay_soft_packed / hamming_soft_packed are __noinline__): the callers sit in unrolled row loops, and four inlined copies
// of the Golay decoder plus three of the Hamming decoder push the soft-decision kernels past the instruction cache (profiles/r01z_imbe_soft_kernel_sass_summary.txt: 33 % of the stall samples were
// instruction fetch).
__device__ __forceinline__ unsigned golay_soft_search(unsigned hard23, const unsigned char* rel, unsigned short* s_cp, unsigned* s_ka,
                                                   unsigned short* s_qa, const DevTables* T, int lane) {
    const unsigned hd = (hard23 >> 11) & 0xfffu;
    const unsigned syn = golay_parity(hd, T) ^ (hard23 & 0x7ffu);
    // five-bit subset sums selected by the lane number: positions 17..21 (high data half), 11..15 (low data half),
    // 1..5 and 6..10 (parity)
    unsigned sA = 0, sB = 0, sL = 0, sH = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const unsigned bit = ((unsigned)lane >> i) & 1u;
        sA += bit * (unsigned)rel[17 + i];
        sB += bit * (unsigned)rel[11 + i];
        sL += bit * (unsigned)rel[1 + i];
        sH += bit * (unsigned)rel[6 + i];
    }
    const unsigned pl = (unsigned)__popc(lane);
    const unsigned hdh = hd >> 6, hdl = hd & 63u;
    __syncwarp();
    s_ka[lane] = (sA << 16) | (pl << 12) | ((((unsigned)lane) ^ hdh) << 6);
    s_qa[lane] = (unsigned short)(((unsigned)T->golay_par_hi[lane] ^ syn) << 1);
    const unsigned kb0 = (sB << 16) | (pl << 12) | (((unsigned)lane) ^ hdl);
    const unsigned kb1 = ((sB + (unsigned)rel[16]) << 16) | ((pl + 1u) << 12) | (((unsigned)lane + 32u) ^ hdl);
    const unsigned pb0 = (unsigned)T->golay_par_lo[lane] << 1;
    const unsigned pb1 = (unsigned)T->golay_par_lo[lane + 32] << 1;
    {
        // cp[64 h + 2 lane + j], j = 0, 1, written as one word per lane and h
        const unsigned pair = sL | ((sL + (unsigned)rel[0]) << 16);
        const unsigned hdup = sH * 0x10001u;
        unsigned* cp32 = reinterpret_cast<unsigned*>(s_cp);
#pragma unroll 8
        for (int h = 0; h < 32; ++h) {
            cp32[32 * h + lane] = pair + __shfl_sync(FULL, hdup, h);
        }
    }
    __syncwarp();
    const unsigned char* cpb = reinterpret_cast<const unsigned char*>(s_cp);
    // The all-ones word is a codeword, so the complement of every candidate is a candidate too, and its key is
    // KMAX - key with no borrow between the fields (cost -> total - cost, differing bits -> 12 - n, data -> 0xfff - data):
    // walk the 32 high halves with a clear top bit, keep the smallest AND the largest key
    const unsigned total = soft_cost_of(rel, 0x7fffffu, 23, lane);
    const unsigned kmax = (total << 16) | (12u << 12) | 0xfffu;
    unsigned best = 0xffffffffu, worst = 0u;
#pragma unroll 8
    for (int a = 0; a < 32; ++a) {
        const unsigned q = s_qa[a];
        const unsigned k = s_ka[a];
        const unsigned c0 = *reinterpret_cast<const unsigned short*>(cpb + (q ^ pb0));
        const unsigned c1 = *reinterpret_cast<const unsigned short*>(cpb + (q ^ pb1));
        const unsigned k0 = ((c0 << 16) + kb0) + k, k1 = ((c1 << 16) + kb1) + k;
        best = min(best, min(k0, k1));
        worst = max(worst, max(k0, k1));
    }
    best = __reduce_min_sync(FULL, min(best, kmax - worst));
    __syncwarp();
    return best;
}

// returns corrected data | changed data bits << 24
__device__ __forceinline__ unsigned golay_soft_body(unsigned hard23, const unsigned char* rel, unsigned short* s_cp, unsigned* s_ka,
                                                   unsigned short* s_qa, const DevTables* T, int lane) {
    int e_hard;
    const unsigned hf = golay_hard(hard23, T, &e_hard);
    const unsigned hd = (hard23 >> 11) & 0xfffu;
    const unsigned syn = golay_parity(hd, T) ^ (hard23 & 0x7ffu);
    // the hard decode's difference pattern (weight <= 3) and its cost
    const unsigned dhf = hf ^ hd;
    const unsigned xhf = (dhf << 11) | (golay_parity(dhf, T) ^ syn);
    const unsigned U = soft_cost_of(rel, xhf, 23, lane);
    if (U == 0u || soft_least_outside(rel, xhf, 23, 7 - __popc(xhf), U, lane) >= U) {
        return hf | ((unsigned)e_hard << 24);
    }
    const unsigned best = golay_soft_search(hard23, rel, s_cp, s_ka, s_qa, T, lane);
    if ((best >> 16) < U) {
        return (best & 0xfffu) | (((best >> 12) & 15u) << 24);
    }
    return hf | ((unsigned)e_hard << 24);
}

__device__ __noinline__ unsigned golay_soft_packed(unsigned hard23, const unsigned char* rel, unsigned short* s_cp, unsigned* s_ka,
                                                   unsigned short* s_qa, const DevTables* T, int lane) {
    return golay_soft_body(hard23, rel, s_cp, s_ka, s_qa, T, lane);
}

// outlined: one shared copy of the decoder (the IMBE front-ends call it four times per frame from unrolled loops);
// the AMBE front-ends, with two calls and smaller kernels, keep it inline (1-2 % faster there)
__device__ __forceinline__ unsigned golay_soft(unsigned hard23, const unsigned char* rel, const SoftScratch& S,
                                               const DevTables* T, int lane, int* errs, bool outlined) {
    const unsigned r = outlined ? golay_soft_packed(hard23, rel, S.cp, S.ka, S.qa, T, lane)
                                : golay_soft_body(hard23, rel, S.cp, S.ka, S.qa, T, lane);
    *errs = (int)(r >> 24);
    return r & 0xffffffu;
}

// Hamming(15,11), both bit layouts (V = 0: ecc.c:366-408, V = 1: the IMBE 7100 variant, ecc.c:422-464).  Same coset
// walk: 11 data bits = 5 high x 6 low, 4 parity bits through a 16-entry key table.  Returns the codeword.
template <int V>
__device__ __forceinline__ unsigned hamming_soft_search(unsigned hard15, const unsigned char* rel, unsigned short* s_cp, unsigned* s_ka,
                                                     unsigned short* s_qa, const DevTables* T, int lane) {
    constexpr int dpos[2][11] = {{2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14}, {4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14}};
    constexpr int ppos[2][4] = {{0, 1, 3, 7}, {0, 1, 2, 3}};
    unsigned dh = 0, hp = 0;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        dh |= ((hard15 >> dpos[V][i]) & 1u) << i;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hp |= ((hard15 >> ppos[V][j]) & 1u) << j;
    }
    const unsigned syn = ((unsigned)T->ham_par_hi[V][dh >> 6] ^ (unsigned)T->ham_par_lo[V][dh & 63u]) ^ hp;
    unsigned sA = 0, sB = 0, sP = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const unsigned bit = ((unsigned)lane >> i) & 1u;
        sA += bit * (unsigned)rel[dpos[V][6 + i]];
        sB += bit * (unsigned)rel[dpos[V][i]];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sP += (((unsigned)lane >> j) & 1u) * (unsigned)rel[ppos[V][j]];
    }
    const unsigned pl = (unsigned)__popc(lane);
    unsigned* cp16 = reinterpret_cast<unsigned*>(s_cp);
    __syncwarp();
    if (lane < 16) {
        cp16[lane] = (sP << 16) | (pl << 11);
    }
    s_ka[lane] = (sA << 16) | (pl << 11) | ((((unsigned)lane) << 6) ^ (dh & 0x7c0u));
    s_qa[lane] = (unsigned short)(((unsigned)T->ham_par_hi[V][lane] ^ syn) << 2);
    const unsigned kb0 = (sB << 16) | (pl << 11) | (((unsigned)lane) ^ (dh & 63u));
    const unsigned kb1 = ((sB + (unsigned)rel[dpos[V][5]]) << 16) | ((pl + 1u) << 11) | (((unsigned)lane + 32u) ^ (dh & 63u));
    const unsigned pb0 = (unsigned)T->ham_par_lo[V][lane] << 2;
    const unsigned pb1 = (unsigned)T->ham_par_lo[V][lane + 32] << 2;
    __syncwarp();
    const unsigned char* cpb = reinterpret_cast<const unsigned char*>(s_cp);
    // all-ones is a codeword of both layouts (every check row has even weight): complements as in golay_soft
    const unsigned total = soft_cost_of(rel, 0x7fffu, 15, lane);
    const unsigned kmax = (total << 16) | (15u << 11) | 0x7ffu;
    unsigned best = 0xffffffffu, worst = 0u;
#pragma unroll 8
    for (int a = 0; a < 16; ++a) {
        const unsigned q = s_qa[a];
        const unsigned k = s_ka[a];
        const unsigned c0 = *reinterpret_cast<const unsigned*>(cpb + (q ^ pb0));
        const unsigned c1 = *reinterpret_cast<const unsigned*>(cpb + (q ^ pb1));
        const unsigned k0 = (c0 + kb0) + k, k1 = (c1 + kb1) + k;
        best = min(best, min(k0, k1));
        worst = max(worst, max(k0, k1));
    }
    best = __reduce_min_sync(FULL, min(best, kmax - worst));
    __syncwarp();
    return best;
}

// returns the codeword | differing bits << 24
template <int V>
__device__ __noinline__ unsigned hamming_soft_packed(unsigned hard15, const unsigned char* rel, unsigned short* s_cp, unsigned* s_ka,
                                                     unsigned short* s_qa, const DevTables* T, int lane) {
    int e_hard;
    const unsigned hf = hamming_hard(hard15, V, T, &e_hard);
    const unsigned xhf = hf ^ hard15;  // at most one position
    const unsigned U = soft_cost_of(rel, xhf, 15, lane);
    if (U == 0u || soft_least_outside(rel, xhf, 15, 3 - __popc(xhf), U, lane) >= U) {
        return hf | ((unsigned)e_hard << 24);
    }
    const unsigned best = hamming_soft_search<V>(hard15, rel, s_cp, s_ka, s_qa, T, lane);
    if ((best >> 16) < U) {
        return (unsigned)T->ham_cw[V][best & 0x7ffu] | (((best >> 11) & 15u) << 24);
    }
    return hf | ((unsigned)e_hard << 24);
}

template <int V>
__device__ __forceinline__ unsigned hamming_soft(unsigned hard15, const unsigned char* rel, const SoftScratch& S,
                                                 const DevTables* T, int lane, int* errs) {
    const unsigned r = hamming_soft_packed<V>(hard15, rel, S.cp, S.ka, S.qa, T, lane);
    *errs = (int)(r >> 24);
    return r & 0xffffffu;
}

// Golay row, hard or soft; `w` holds the 23 received bits, returns the row with corrected data bits and
// the received parity bits (both decoders echo the input parity, ecc.c:290-292,352-355).
__device__ __forceinline__ unsigned golay_row(unsigned w, const unsigned char* rel, int soft, const SoftScratch& S,
                                              const DevTables* T, int lane, int* errs, bool outlined = true) {
    unsigned data = soft ? golay_soft(w, rel, S, T, lane, errs, outlined) : golay_hard(w, T, errs);
    return (data << 11) | (w & 0x7ffu);
}

template <int V>
__device__ __forceinline__ unsigned hamming_row(unsigned w, const unsigned char* rel, int soft, const SoftScratch& S,
                                                const DevTables* T, int lane, int* errs) {
    return soft ? hamming_soft<V>(w, rel, S, T, lane, errs) : hamming_hard(w, V, T, errs);
}

__device__ __forceinline__ unsigned pn_bit(unsigned p0, int k, const DevTables* T) {
    return (((unsigned)T->pnA[k] * p0 + (unsigned)T->pnC[k]) & 0xffffu) >> 15;
}

__device__ __forceinline__ unsigned getbit(const unsigned dw[3], int i) {
    // (selects, not an indexed load: a run-time index would put the three words into local memory)
    const unsigned w = (i < 32) ? dw[0] : (i < 64 ? dw[1] : dw[2]);
    return (w >> (i & 31)) & 1u;
}

// ---- the channel front-end in its four steps: read | C0 | de-scramble | data ----------------------------------------
// (the reference exposes them one by one: mbe_ecc<Codec>C0, mbe_demodulate<Codec>Data, mbe_ecc<Codec>Data,
//  mbe_convertImbe7100to7200; the frame kernels run them back to back on rows that stay in registers.)
//   ws_rel  : per-warp scratch for reliabilities (8*24 bytes)
//   rb      : per-warp scratch, 8 words (corrected rows, so that lanes can index them dynamically)
//   S       : per-warp soft-decision scratch (soft only)
//   packed  : hard bits packed eight per byte, MSB first, in the row-major order of the reference's fr[rows][cols] or in
//             the transmission order given by the context's channel map
__device__ __forceinline__ int fe_rows(int codec) { return (codec == MBE_B200_IMBE7200X4400) ? 8 : (codec == MBE_B200_IMBE7100X4400 ? 7 : 4); }
__device__ __forceinline__ int fe_cols(int codec) { return (codec == MBE_B200_IMBE7200X4400) ? 23 : 24; }

// rows of the frame as ballot words (lane j = column j); returns true when a bit is neither 0 nor 1
__device__ __forceinline__ bool fe_read(int codec, int soft, int packed, const uint8_t* __restrict__ fr, unsigned row[8],
                                        unsigned char* ws_rel, const DevTables* T, int lane) {
    const int rows = fe_rows(codec), cols = fe_cols(codec);
    bool bad = false;
    __syncwarp();  // the scratch may alias rows the previous frame's stores have just read
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        unsigned b = 0;
        if (r < rows && lane < cols) {
            int idx = r * cols + lane;
            unsigned v;
            if (soft) {
                v = fr[2 * idx];
                ws_rel[r * 24 + lane] = fr[2 * idx + 1];
            } else if (packed) {
                const unsigned src = T->chan_src[codec][idx];  // transmitted bit that lands here (identity by default)
                v = (src == 0xffffu) ? 0u : (((unsigned)fr[src >> 3] >> (7 - (src & 7))) & 1u);
            } else {
                v = fr[idx];
            }
            bad |= (v > 1u);
            b = v & 1u;
        }
        row[r] = __ballot_sync(FULL, b);
    }
    __syncwarp();
    return __any_sync(FULL, bad);
}

// C0 in place; returns its error count
__device__ __forceinline__ int fe_c0(int codec, int soft, unsigned row[8], unsigned char* ws_rel, const SoftScratch& S,
                                     const DevTables* T, int lane) {
    int c0 = 0;
    if (codec == MBE_B200_IMBE7200X4400) {
        row[0] = golay_row(row[0], ws_rel, soft, S, T, lane, &c0);
    } else if (codec == MBE_B200_IMBE7100X4400) {
        // C0: columns 1..18 zero-extended to a 23-bit Golay word; pad bits are fully reliable zeros
        if (soft) {
            // build a dedicated reliability vector for the padded word in the row-7 scratch slot
            if (lane < 23) {
                ws_rel[7 * 24 + lane] = (lane < 18) ? ws_rel[lane + 1] : (unsigned char)255;
            }
            __syncwarp();
        }
        unsigned w0 = (row[0] >> 1) & 0x3ffffu;
        unsigned d0 = golay_row(w0, ws_rel + 7 * 24, soft, S, T, lane, &c0);
        row[0] = (row[0] & ~(0x3ffffu << 1)) | ((d0 & 0x3ffffu) << 1);
    } else {
        // AMBE 3600: C0 = Golay on columns 1..23 + overall parity in column 0
        unsigned g = golay_row((row[0] >> 1) & 0x7fffffu, ws_rel + 1, soft, S, T, lane, &c0, false);
        row[0] = (row[0] & 1u) | (g << 1);
        if (c0 == 0 && (__popc(row[0] & 0xffffffu) & 1)) {
            row[0] ^= 1u;
            c0 = 1;
        }
    }
    return c0;
}

// PN de-scrambling of the protected rows with the sequence seeded by C0's data bits
__device__ __forceinline__ void fe_demod(int codec, unsigned row[8], const DevTables* T, int lane) {
    if (codec == MBE_B200_IMBE7200X4400) {
        const unsigned p0 = (16u * ((row[0] >> 11) & 0xfffu)) & 0xffffu;
        // PN masks: rows 1..3 use k = 1 + 23 (r-1) + (22 - j); rows 4..6 use k = 70 + 15 (r-4) + (14 - j)
#pragma unroll
        for (int r = 1; r < 4; ++r) {
            unsigned b = (lane < 23) ? pn_bit(p0, 1 + 23 * (r - 1) + (22 - lane), T) : 0u;
            row[r] ^= __ballot_sync(FULL, b);
        }
#pragma unroll
        for (int r = 4; r < 7; ++r) {
            unsigned b = (lane < 15) ? pn_bit(p0, 70 + 15 * (r - 4) + (14 - lane), T) : 0u;
            row[r] ^= __ballot_sync(FULL, b);
        }
    } else if (codec == MBE_B200_IMBE7100X4400) {
        const unsigned seed = (row[0] >> 12) & 0x7fu;
        const unsigned p0 = (16u * seed) & 0xffffu;
        {
            unsigned b = (lane < 24) ? pn_bit(p0, 1 + (23 - lane), T) : 0u;
            row[1] ^= __ballot_sync(FULL, b);
        }
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            unsigned b = (lane < 23) ? pn_bit(p0, 25 + 23 * (r - 2) + (22 - lane), T) : 0u;
            row[r] ^= __ballot_sync(FULL, b);
        }
#pragma unroll
        for (int r = 4; r < 6; ++r) {
            unsigned b = (lane < 15) ? pn_bit(p0, 71 + 15 * (r - 4) + (14 - lane), T) : 0u;
            row[r] ^= __ballot_sync(FULL, b);
        }
    } else {
        const unsigned p0 = (16u * ((row[0] >> 12) & 0xfffu)) & 0xffffu;
        unsigned b = (lane < 23) ? pn_bit(p0, 1 + (22 - lane), T) : 0u;
        row[1] ^= __ballot_sync(FULL, b);
    }
}

// K-dependent permutation of the 88 IMBE 7100 parameter bits to the 7200 layout (imbe7100x4400.c:380-437)
__device__ __forceinline__ void fe_convert7100(const unsigned pre[3], unsigned dw[3], const DevTables* T, int lane) {
    unsigned b0 = 0;
    {
        const int idx[8] = {1, 2, 3, 4, 5, 6, 86, 87};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            b0 = (b0 << 1) | getbit(pre, idx[i]);
        }
    }
    const int K = (int)T->imbe_K[b0];
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        int j = 32 * w + lane;
        unsigned b = 0;
        if (j < 88) {
            int src;
            if (j == 87) {
                src = 0;
            } else if (j < 48) {
                src = (j + 1 <= 41) ? j + 1 : j + K + 3;
            } else if (j < 48 + K) {
                src = 44 + (j - 48);
            } else if (j == 48 + K) {
                src = 42;
            } else if (j == 49 + K) {
                src = 43;
            } else {
                int n = j - K - 2;
                src = (n + 1 <= 41) ? n + 1 : n + K + 3;
            }
            b = getbit(pre, src);
        }
        dw[w] = __ballot_sync(FULL, b);
    }
}

// ECC of the remaining rows and packing of the parameter bits as three ballot words (bit i of the reference's
// imbe_d/ambe_d = bit (i & 31) of dw[i >> 5]); returns the protected-field error count.  convert: IMBE 7100 only, also
// apply fe_convert7100 (the frame paths do; mbe_eccImbe7100x4400Data alone does not).
__device__ __forceinline__ int fe_data(int codec, int soft, bool convert, unsigned row[8], unsigned char* ws_rel,
                                       const SoftScratch& S, unsigned* rb, const DevTables* T, int lane, unsigned dw[3],
                                       int* c4_out) {
    int prot = 0, c4 = 0, e;
    if (codec == MBE_B200_IMBE7200X4400) {
#pragma unroll
        for (int r = 1; r < 4; ++r) {
            row[r] = golay_row(row[r] & 0x7fffffu, ws_rel + 24 * r, soft, S, T, lane, &e);
            prot += e;
        }
#pragma unroll
        for (int r = 4; r < 7; ++r) {
            row[r] = hamming_row<0>(row[r] & 0x7fffu, ws_rel + 24 * r, soft, S, T, lane, &e);
            prot += e;
            if (r == 4) {
                c4 = e;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                rb[r] = row[r];
            }
        }
        __syncwarp();
        // pack: 4 x 12 Golay data bits MSB first, 3 x 11 Hamming bits 14..4, 7 raw bits 6..0
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            int o = 32 * w + lane;
            unsigned b = 0;
            if (o < 48) {
                b = (rb[o / 12] >> (22 - (o % 12))) & 1u;
            } else if (o < 81) {
                int q = o - 48;
                b = (rb[4 + q / 11] >> (14 - (q % 11))) & 1u;
            } else if (o < 88) {
                b = (row[7] >> (6 - (o - 81))) & 1u;
            }
            dw[w] = __ballot_sync(FULL, b);
        }
    } else if (codec == MBE_B200_IMBE7100X4400) {
        // row 1 carries its Golay word in columns 1..23
        {
            unsigned g = golay_row((row[1] >> 1) & 0x7fffffu, ws_rel + 24 * 1 + 1, soft, S, T, lane, &e);
            row[1] = (row[1] & 1u) | (g << 1);
            prot = e;
        }
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            row[r] = golay_row(row[r] & 0x7fffffu, ws_rel + 24 * r, soft, S, T, lane, &e);
            prot += e;
        }
#pragma unroll
        for (int r = 4; r < 6; ++r) {
            row[r] = hamming_row<1>(row[r] & 0x7fffu, ws_rel + 24 * r, soft, S, T, lane, &e);
            prot += e;
            if (r == 4) {
                c4 = e;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                rb[r] = row[r];
            }
        }
        __syncwarp();
        // 7100 layout: 7 bits of row 0 (18..12), row 1 cols 23..12, rows 2,3 bits 22..11, rows 4,5 bits 14..4,
        // row 6 bits 22..0
        unsigned pre[3];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            int o = 32 * w + lane;
            unsigned b = 0;
            if (o < 7) {
                b = (row[0] >> (18 - o)) & 1u;
            } else if (o < 19) {
                b = (row[1] >> (23 - (o - 7))) & 1u;
            } else if (o < 43) {
                int q = o - 19;
                b = (rb[2 + q / 12] >> (22 - (q % 12))) & 1u;
            } else if (o < 65) {
                int q = o - 43;
                b = (rb[4 + q / 11] >> (14 - (q % 11))) & 1u;
            } else if (o < 88) {
                b = (row[6] >> (22 - (o - 65))) & 1u;
            }
            pre[w] = __ballot_sync(FULL, b);
        }
        if (convert) {
            fe_convert7100(pre, dw, T, lane);
        } else {
            dw[0] = pre[0];
            dw[1] = pre[1];
            dw[2] = pre[2];
        }
    } else {
        row[1] = (row[1] & 0x800000u) | golay_row(row[1] & 0x7fffffu, ws_rel + 24, soft, S, T, lane, &prot, false);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            int o = 32 * w + lane;
            unsigned b = 0;
            if (o < 12) {
                b = (row[0] >> (23 - o)) & 1u;
            } else if (o < 24) {
                b = (row[1] >> (22 - (o - 12))) & 1u;
            } else if (o < 35) {
                b = (row[2] >> (10 - (o - 24))) & 1u;
            } else if (o < 49) {
                b = (row[3] >> (13 - (o - 35))) & 1u;
            }
            dw[w] = __ballot_sync(FULL, b);
        }
        dw[2] = 0;
    }
    *c4_out = c4;
    return prot;
}

// Decode one frame held in global memory: the four steps back to back.
__device__ __forceinline__ FrontResult front_end(int codec, int soft, int packed, const uint8_t* __restrict__ fr, unsigned dw[3],
                                                 unsigned char* ws_rel, const SoftScratch& S, unsigned* rb,
                                                 const DevTables* T, int lane) {
    FrontResult R;
    unsigned row[8];
    if (fe_read(codec, soft, packed, fr, row, ws_rel, T, lane)) {
        R.status = -2;  // MBE_STATUS_INVALID_BITS
        R.c0 = R.prot = R.c4 = 0;
        R.flags = 0;
        dw[0] = dw[1] = dw[2] = 0;
        return R;
    }
    int c4 = 0;
    const int c0 = fe_c0(codec, soft, row, ws_rel, S, T, lane);
    fe_demod(codec, row, T, lane);
    const int prot = fe_data(codec, soft, true, row, ws_rel, S, rb, T, lane, dw, &c4);
    R.flags = (codec <= MBE_B200_IMBE7100X4400) ? (0x0002u | 0x0004u) : 0x0002u;
    if (soft) {
        R.flags |= 0x0001u;
    }
    R.c0 = c0;
    R.prot = prot;
    R.c4 = c4;
    R.status = c0 + prot;
    return R;
}

}  // namespace mbe
