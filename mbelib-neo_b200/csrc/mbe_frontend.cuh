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

// Soft-decision decode of one row: exhaustive maximum-likelihood search over all codewords with the
// reference's tie-break order (src/ecc/ecc.c:54-67): lowest cost, then the candidate equal to the
// hard decode, then fewest differing bits, then lowest data index.  Cost = sum of reliabilities of
// the positions where the candidate differs from the received hard bits, evaluated through three
// byte-indexed partial-sum tables built per row in shared memory.
//   cost_tab: 640 uint16 of per-warp scratch.
__device__ __forceinline__ void soft_build_cost(unsigned short* cost_tab, const unsigned char* rel, int nbits, int lane) {
    // tab[g][v] = sum over set bits b of v of rel[8g + b]
    for (int e = lane; e < 640; e += 32) {
        int g = e >> 8, v = e & 255;
        if (e >= 512) {
            g = 2;
            v = e - 512;
        }
        int s = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            int pos = 8 * g + b;
            if (((v >> b) & 1) && pos < nbits) {
                s += (int)rel[pos];
            }
        }
        cost_tab[e] = (unsigned short)s;
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned golay_soft(unsigned hard23, const unsigned char* rel, unsigned short* cost_tab,
                                               const DevTables* T, int lane, int* errs) {
    int dummy;
    const unsigned hard_fixed = golay_hard(hard23, T, &dummy);
    soft_build_cost(cost_tab, rel, 23, lane);
    unsigned best = 0xffffffffu;
    for (unsigned data = (unsigned)lane; data < 4096u; data += 32u) {
        unsigned x = T->golay_cw[data] ^ hard23;
        unsigned cost = (unsigned)cost_tab[x & 255u] + (unsigned)cost_tab[256 + ((x >> 8) & 255u)]
                        + (unsigned)cost_tab[512 + (x >> 16)];
        unsigned key = (cost << 17) | ((data != hard_fixed) ? (1u << 16) : 0u) | ((unsigned)__popc(x >> 11) << 12) | data;
        best = min(best, key);
    }
    best = __reduce_min_sync(FULL, best);
    __syncwarp();
    *errs = (int)((best >> 12) & 15u);
    return best & 0xfffu;
}

__device__ __forceinline__ unsigned hamming_soft(unsigned hard15, int variant, const unsigned char* rel,
                                                 unsigned short* cost_tab, const DevTables* T, int lane, int* errs) {
    int dummy;
    const unsigned hard_fixed = hamming_hard(hard15, variant, T, &dummy);
    soft_build_cost(cost_tab, rel, 15, lane);
    unsigned best = 0xffffffffu;
    for (unsigned data = (unsigned)lane; data < 2048u; data += 32u) {
        unsigned cw = (unsigned)T->ham_cw[variant][data];
        unsigned x = cw ^ hard15;
        unsigned cost = (unsigned)cost_tab[x & 255u] + (unsigned)cost_tab[256 + (x >> 8)];
        unsigned key = (cost << 16) | ((cw != hard_fixed) ? (1u << 15) : 0u) | ((unsigned)__popc(x) << 11) | data;
        best = min(best, key);
    }
    best = __reduce_min_sync(FULL, best);
    __syncwarp();
    *errs = (int)((best >> 11) & 15u);
    return (unsigned)T->ham_cw[variant][best & 0x7ffu];
}

// Golay row, hard or soft; `w` holds the 23 received bits, returns the row with corrected data bits and
// the received parity bits (both decoders echo the input parity, ecc.c:290-292,352-355).
__device__ __forceinline__ unsigned golay_row(unsigned w, const unsigned char* rel, int soft, unsigned short* cost_tab,
                                              const DevTables* T, int lane, int* errs) {
    unsigned data = soft ? golay_soft(w, rel, cost_tab, T, lane, errs) : golay_hard(w, T, errs);
    return (data << 11) | (w & 0x7ffu);
}

__device__ __forceinline__ unsigned hamming_row(unsigned w, int variant, const unsigned char* rel, int soft,
                                                unsigned short* cost_tab, const DevTables* T, int lane, int* errs) {
    return soft ? hamming_soft(w, variant, rel, cost_tab, T, lane, errs) : hamming_hard(w, variant, T, errs);
}

__device__ __forceinline__ unsigned pn_bit(unsigned p0, int k, const DevTables* T) {
    return (((unsigned)T->pnA[k] * p0 + (unsigned)T->pnC[k]) & 0xffffu) >> 15;
}

__device__ __forceinline__ unsigned getbit(const unsigned dw[3], int i) {
    return (dw[i >> 5] >> (i & 31)) & 1u;
}

// Decode one frame held in global memory.  Outputs the parameter bits as three ballot words
// (bit i of the reference's imbe_d/ambe_d = bit (i & 31) of dw[i >> 5]).
//   ws_rel  : per-warp scratch for reliabilities (8*24 bytes)
//   rb      : per-warp scratch, 8 words (corrected rows, so that lanes can index them dynamically)
//   cost_tab: per-warp scratch, 640 uint16 (soft only)
//   packed  : hard bits packed eight per byte, MSB first, in the row-major order of the reference's fr[rows][cols]
__device__ __forceinline__ FrontResult front_end(int codec, int soft, int packed, const uint8_t* __restrict__ fr, unsigned dw[3],
                                                 unsigned char* ws_rel, unsigned short* cost_tab, unsigned* rb,
                                                 const DevTables* T, int lane) {
    FrontResult R;
    const int rows = (codec == MBE_B200_IMBE7200X4400) ? 8 : (codec == MBE_B200_IMBE7100X4400 ? 7 : 4);
    const int cols = (codec == MBE_B200_IMBE7200X4400) ? 23 : 24;
    unsigned row[8];
    bool bad = false;
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
                v = ((unsigned)fr[idx >> 3] >> (7 - (idx & 7))) & 1u;
            } else {
                v = fr[idx];
            }
            bad |= (v > 1u);
            b = v & 1u;
        }
        row[r] = __ballot_sync(FULL, b);
    }
    __syncwarp();
    if (__any_sync(FULL, bad)) {
        R.status = -2;  // MBE_STATUS_INVALID_BITS
        R.c0 = R.prot = R.c4 = 0;
        R.flags = 0;
        dw[0] = dw[1] = dw[2] = 0;
        return R;
    }
    int c0 = 0, prot = 0, c4 = 0, e;

    if (codec == MBE_B200_IMBE7200X4400) {
        row[0] = golay_row(row[0], ws_rel, soft, cost_tab, T, lane, &c0);
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
#pragma unroll
        for (int r = 1; r < 4; ++r) {
            row[r] = golay_row(row[r] & 0x7fffffu, ws_rel + 24 * r, soft, cost_tab, T, lane, &e);
            prot += e;
        }
#pragma unroll
        for (int r = 4; r < 7; ++r) {
            row[r] = hamming_row(row[r] & 0x7fffu, 0, ws_rel + 24 * r, soft, cost_tab, T, lane, &e);
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
        R.flags = 0x0002u | 0x0004u;
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
        unsigned d0 = golay_row(w0, ws_rel + 7 * 24, soft, cost_tab, T, lane, &c0);
        row[0] = (row[0] & ~(0x3ffffu << 1)) | ((d0 & 0x3ffffu) << 1);
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
        // row 1 carries its Golay word in columns 1..23
        {
            unsigned g = golay_row((row[1] >> 1) & 0x7fffffu, ws_rel + 24 * 1 + 1, soft, cost_tab, T, lane, &e);
            row[1] = (row[1] & 1u) | (g << 1);
            prot = e;
        }
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            row[r] = golay_row(row[r] & 0x7fffffu, ws_rel + 24 * r, soft, cost_tab, T, lane, &e);
            prot += e;
        }
#pragma unroll
        for (int r = 4; r < 6; ++r) {
            row[r] = hamming_row(row[r] & 0x7fffu, 1, ws_rel + 24 * r, soft, cost_tab, T, lane, &e);
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
        // permutation to the 7200 layout (imbe7100x4400.c:380-437)
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
        R.flags = 0x0002u | 0x0004u;
    } else {
        // AMBE 3600: C0 = Golay on columns 1..23 + overall parity in column 0
        unsigned g = golay_row((row[0] >> 1) & 0x7fffffu, ws_rel + 1, soft, cost_tab, T, lane, &c0);
        row[0] = (row[0] & 1u) | (g << 1);
        if (c0 == 0 && (__popc(row[0] & 0xffffffu) & 1)) {
            row[0] ^= 1u;
            c0 = 1;
        }
        const unsigned p0 = (16u * ((row[0] >> 12) & 0xfffu)) & 0xffffu;
        {
            unsigned b = (lane < 23) ? pn_bit(p0, 1 + (22 - lane), T) : 0u;
            row[1] ^= __ballot_sync(FULL, b);
        }
        row[1] = (row[1] & 0x800000u) | golay_row(row[1] & 0x7fffffu, ws_rel + 24, soft, cost_tab, T, lane, &prot);
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
        R.flags = 0x0002u;
    }
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
