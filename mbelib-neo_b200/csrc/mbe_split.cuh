// The frame program as THREE kernels:
//
//   parameter kernel   mbe_stream_kernel<CODEC, SOFT, MODE, /*SPLIT=*/true>: everything that is serial per stream - ECC,
//                      parameter decode, state machine, enhancement, smoothing, phase update, noise-generator state, the
//                      start state of every oscillator - and the frames that need no speech synthesis (silence, tones,
//                      comfort noise).  For every frame it writes a DESCRIPTOR: what the synthesis needs and nothing else.
//                      Its blocks walk their streams in lockstep like the fused kernel's (the program is ~85 KB of SASS;
//                      warps on the same code share their instruction fetches).
//   bank kernel        mbe_split_bank_kernel: the voiced oscillator bank (mbelib.c:953-1040) and nothing else - a loop of a
//                      few hundred instructions that stays in the instruction cache, so its warps need no lockstep: a warp
//                      pools the components of BG consecutive frames into 32-slot passes over its private 32 x 32 tile
//                      (phase A: lane = slot runs the reference's unfused rotation recurrence; phase B: lane = sample adds
//                      each frame's slots in the reference's order).  One warp's latency-bound ordered sums overlap another
//                      warp's FP32-dense recurrences, and a frame with many harmonics delays nobody else.
//   unvoiced kernel    mbe_split_unvoiced_kernel: 256-point FFT / IFFT unvoiced synthesis with weighted overlap-add
//                      (mbe_unvoiced_fft.c:714-761), soft clip, float -> int16.  It owns the previousUw arrays of the
//                      stream's three parameter sets and replays the struct copies the state machine asked for (4-bit op
//                      codes in the descriptor).  Every frame costs the same here, so its blocks run in lockstep for free.
//
// Why three: a first two-kernel cut (one synthesis kernel with decoupled warps, 59 KB of SASS) spent 51 % of its issue
// slots waiting for instructions (profiles/experiments/r02i_*): warps that are not on the same code cannot share fetches,
// and the straight-line transforms are far larger than the 32 KB L1.5 instruction cache.
#pragma once
#include "mbe_common.cuh"
#include "mbe_synth.cuh"

namespace mbe {

// ---- frame descriptor (words) --------------------------------------------------------------------------------------
constexpr int D_OPS = 0;     // previousUw ops of the frame, 4 bits each, in order (OP_*; OP_SYNTH = the synthesis itself)
constexpr int D_INFO = 1;    // bit 0: frame runs the synthesis; bits 8..15 cur L; bits 16..23 oscillator slots (components)
constexpr int D_CW0 = 2;     // cur_mp->w0
constexpr int D_DW0 = 3;     // cur_mp->w0 - prev_mp_enhanced->w0 (the chirp term of the interpolated harmonics)
constexpr int D_SEED = 4;    // cur_mp->noiseSeed before the frame's noise buffer is built (< 0: cold start)
constexpr int D_CU = 6;      // 2 words: bit l-1 = (cur Vl[l] == 0), l = 1..56, after smoothing
constexpr int D_CML = 16;    // cur Ml[1..56]
constexpr int D_OV = 72;     // cur_mp->noiseOverlap[96] before the frame
constexpr int D_KIND = 168;  // 112 bytes: per slot harmonic << 2 | kind (0 previous-frame window, 1 current-frame window,
                             // 2 phase / amplitude interpolated), list order = the reference's summation order
constexpr int D_G = 196;     // per slot, 112 words each: gain            | interpolated: phase increment per sample
constexpr int D_C = 308;     //                           cos(start phase) |               prev PHIl[l]
constexpr int D_S = 420;     //                           sin(start phase) |               prev Ml[l]
constexpr int D_CD = 532;    //                           cos(l w0)        |               cur Ml[l]
constexpr int D_SD = 644;    //                           sin(l w0)
constexpr int D_VOICED = D_G;  // the bank kernel leaves the frame's 160 voiced samples here (its slot records are dead by then)
constexpr int DESC_WORDS = 768;
static_assert(D_SD + 112 <= DESC_WORDS && D_VOICED + NS <= D_S, "descriptor layout");

#ifndef MBE_BG
#define MBE_BG 3
#endif
#ifndef MBE_BWARPS
#define MBE_BWARPS 8
#endif
#ifndef MBE_BMINB
#define MBE_BMINB 4
#endif
#ifndef MBE_UWARPS
#define MBE_UWARPS 8
#endif
#ifndef MBE_UMINB
#define MBE_UMINB 4
#endif
#ifndef MBE_U_LOCKSTEP
#define MBE_U_LOCKSTEP 0
#endif
constexpr int BG = MBE_BG;            // frames pooled per warp of the bank kernel
constexpr int B_WARPS = MBE_BWARPS;   // warps per block of the bank kernel
constexpr int B_MINB = MBE_BMINB;
constexpr int U_WARPS = MBE_UWARPS;   // warps (= streams) per block of the unvoiced kernel
constexpr int U_MINB = MBE_UMINB;

// ---- parameter-kernel side -----------------------------------------------------------------------------------------
// make_noise without the FFT input: advances the generator (seed, overlap of cur_mp and prev_mp_enhanced) exactly like
// make_noise and hands back what the synthesis kernel needs to rebuild the frame's buffer: the overlap and the seed as they
// were before (mbe_unvoiced_fft.c:304-341)
__device__ __forceinline__ void make_noise_state(WarpWS& ws, float* cur_overlap, float* enh_overlap, const BlockTables* bt,
                                                 int lane, float ov[3], float* seed_before) {
    ParmsSmall& cur = ws.cur;
    const float seed = cur.noiseSeed;
    *seed_before = seed;
    if (seed < 0.0f) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            ov[r] = 0.0f;
            cur_overlap[32 * r + lane] = 0.0f;
            enh_overlap[32 * r + lane] = 0.0f;
        }
        const float ns = ws.rng.uv_override ? (float)ws.rng.uv_seed : 3147.0f;
        __syncwarp();
        if (lane == 0) {
            cur.noiseSeed = ns;
            ws.rng.uv_override = 0;
        }
        __syncwarp();
        return;
    }
    constexpr unsigned A32 = lcg_pow_a(32), C32 = lcg_pow_c(32);
    constexpr unsigned A64 = lcg_pow_a(64), C64 = lcg_pow_c(64);
    unsigned sc = ((unsigned)seed) % 53125u;
    const uint2 jl = bt->uv_jump[lane];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ov[r] = cur_overlap[32 * r + lane];
    }
    sc = (A64 * sc + C64) % 53125u;   // the first 64 new samples never reach the overlap
#pragma unroll
    for (int c = 2; c < 5; ++c) {
        const unsigned st = (jl.x * sc + jl.y) % 53125u;
        sc = (A32 * sc + C32) % 53125u;
        const float v = (float)st;
        cur_overlap[32 * (c - 2) + lane] = v;
        enh_overlap[32 * (c - 2) + lane] = v;
    }
    const unsigned stn = sc;
    __syncwarp();
    if (lane == 0) {
        cur.noiseSeed = (float)stn;
    }
    __syncwarp();
}

// the arrays of a synthesis frame's descriptor (after synth_begin: phases updated, lengths reconciled, component list
// built): what the unvoiced kernel reads, and one record per oscillator slot - the start state the fused kernel's bank
// computes per lane (voiced_bank_block), here once, by the kernel that runs in lockstep anyway
__device__ __forceinline__ void emit_desc_arrays(uint32_t* d, const WarpWS& ws, const float ov[3], const DevTables* T,
                                                 int lane) {
    const ParmsSmall& cur = ws.cur;
    const EnhSmall& prev = ws.enh;
    unsigned cu[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int l = 1 + lane + 32 * r;
        int cvl = 2;
        if (l <= 56) {
            cvl = cur.Vl[l];
            d[D_CML + l - 1] = __float_as_uint(cur.Ml[l]);
        }
        cu[r] = __ballot_sync(FULL, cvl == 0);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[D_OV + 32 * r + lane] = __float_as_uint(ov[r]);
    }
    if (lane < 2) {
        d[D_CU + lane] = lane ? cu[1] : cu[0];
    }
    const int ncomp = ws.ncomp;
    const float cw0 = cur.w0, pw0 = prev.w0;
    unsigned char* kinds = reinterpret_cast<unsigned char*>(d + D_KIND);
#pragma unroll 1
    for (int k = lane; k < ncomp; k += 32) {
        const int id = ws.comp[k];
        const int l = id >> 2;
        float g, c, s, cd, sd = 0.0f;
        if ((id & 3) == 2) {
            const float pw0l = pw0 * (float)l;
            const float dphi = cur.PHIl[l] - prev.PHIl[l] - (((pw0 + cw0) * (float)(l * NS)) / 2.0f);
            const float dw = (1.0f / (float)NS) * (dphi - (2.0f * MBE_PI_F * floorf((dphi + MBE_PI_F) / (2.0f * MBE_PI_F))));
            g = pw0l + dw;
            c = prev.PHIl[l];
            s = prev.Ml[l];
            cd = cur.Ml[l];
        } else {
            float step, ph, w0;
            int row;
            if ((id & 3) == 0) {
                w0 = pw0;
                row = ws.w0row_enh;
                step = w0 * (float)l;
                ph = prev.PHIl[l];
                g = 2.0f * prev.Ml[l];
            } else {
                w0 = cw0;
                row = ws.w0row_prev;   // (after the enhancement: the row that matched this frame's fundamental)
                step = w0 * (float)l;
                ph = cur.PHIl[l] - (step * (float)NS);
                g = 2.0f * cur.Ml[l];
            }
            const bool rowok = row >= 0 && row < COSW_ROWS;
            const int rr = rowok ? row : 0;
            const float tw0 = T->cosw_w0[rr];
            float2 dd = T->stepsc[rr][l];
            const float2 p = dev_sincosf(ph);
            if (!(rowok && __float_as_uint(tw0) == __float_as_uint(w0))) {
                dd = dev_sincosf(step);
            }
            sd = dd.x;
            cd = dd.y;
            s = p.x;
            c = p.y;
        }
        kinds[k] = (unsigned char)id;
        d[D_G + k] = __float_as_uint(g);
        d[D_C + k] = __float_as_uint(c);
        d[D_S + k] = __float_as_uint(s);
        d[D_CD + k] = __float_as_uint(cd);
        d[D_SD + k] = __float_as_uint(sd);
    }
    __syncwarp();   // the frame's hand-over rewrites ws.enh: every lane must be done reading it
}

// the descriptor's header; every frame of the launch gets one (a frame without synthesis may still move previousUw).
// emit_desc_info runs BEFORE the frame's hand-over (prev_mp_enhanced <- cur_mp changes w0), emit_desc_ops after it (the
// hand-over may record one more op).
__device__ __forceinline__ void emit_desc_info(uint32_t* d, const WarpWS& ws, int go, float seed_before, int lane) {
    if (lane >= 1 && lane < 5) {
        unsigned v;
        if (lane == 1) {
            v = (go ? 1u : 0u) | ((unsigned)(ws.cur.L & 255) << 8) | ((unsigned)(ws.ncomp & 255) << 16);
        } else if (lane == 2) {
            v = __float_as_uint(ws.cur.w0);
        } else if (lane == 3) {
            v = __float_as_uint(ws.cur.w0 - ws.enh.w0);
        } else {
            v = __float_as_uint(seed_before);
        }
        d[lane] = v;
    }
    __syncwarp();   // (the hand-over that follows rewrites ws.enh)
}
__device__ __forceinline__ void emit_desc_ops(uint32_t* d, const WarpWS& ws, int lane) {
    if (lane == 0) {
        d[D_OPS] = ws.ops;
    }
}

// ---- bank kernel -----------------------------------------------------------------------------------------------------
struct SynthArgs {
    int first_stream, io_base, n_streams, n_frames;
    uint32_t* desc;         // [n_streams][n_frames][DESC_WORDS]
    int16_t* pcm;           // [..][n_frames][160] (frames without synthesis were written by the parameter kernel)
    float* pcmf;
    float pcmf_scale;
    uint32_t* state;        // [max_streams][STATE_WORDS]: only the previousUw words are touched
    const DevTables* tab;
    unsigned long long* counters;   // profiling (may be null): [0] oscillator slots run, [1] of them interpolated harmonics,
                                    // [2] frames synthesised
};

// Oscillator tile of the bank kernel: 32 samples x 32 slots, rows padded to 36 words: phase A (lane = slot, one row per
// store) and phase B (lane = sample, LDS.128 over four consecutive slots) are both bank-conflict free, and every access is
// a base register + immediate offset (no swizzle arithmetic in the loops)
constexpr int BT_STRIDE = 36;
#ifndef MBE_BANK_STEPS
#define MBE_BANK_STEPS 16
#endif
constexpr int BANK_STEPS = MBE_BANK_STEPS;   // oscillator steps per loop body of phase A (8 / 16 / 32)
struct __align__(16) BankWS {
    float tile[32 * BT_STRIDE];
    float out[BG][NS];      // voiced samples of the group's frames (lane i owns i, 32 + i, ...)
};

// smem: voiced window halves (BlockTables::voiced_win layout, 336 floats) | n / 160 for n = 0..159 | per-warp workspaces
constexpr int BANK_TAB_FLOATS = 336 + NS;

__global__ void __launch_bounds__(B_WARPS * 32, B_MINB) mbe_split_bank_kernel(const SynthArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* vwin = reinterpret_cast<float*>(smem_raw);
    float* nfrac = vwin + 336;   // (float)n / (float)N of the amplitude interpolation (mbelib.c:963), the same division
    BankWS* wsa = reinterpret_cast<BankWS*>(smem_raw + BANK_TAB_FLOATS * sizeof(float));
    for (int i = threadIdx.x; i < 2 * NS; i += blockDim.x) {
        vwin[i < NS ? i : i + (WIN_PREV - NS)] = A.tab->voiced_win[i];
    }
    for (int i = threadIdx.x; i < NS; i += blockDim.x) {
        nfrac[i] = (float)i / (float)NS;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    BankWS& ws = wsa[warp];
    float* tile = ws.tile;
    const long long n_items = (long long)A.n_streams * A.n_frames;
    const long long n_groups = (n_items + BG - 1) / BG;
    const long long warps_total = (long long)gridDim.x * B_WARPS;
    // the grid is resident (a few blocks per SM); every warp walks groups of BG consecutive frames with a grid stride
#pragma unroll 1
    for (long long grp = (long long)blockIdx.x * B_WARPS + warp; grp < n_groups; grp += warps_total) {
        const long long i0 = grp * BG;
        const int n_live = (int)min((long long)BG, n_items - i0);
        uint32_t* const d0 = A.desc + (size_t)i0 * DESC_WORDS;

        int off[BG + 1], cnt[BG];
        int total = 0;
        unsigned go_mask = 0;   // frames of the group that run the synthesis (bit q)
#pragma unroll
        for (int q = 0; q < BG; ++q) {
            off[q] = total;
            cnt[q] = 0;
            if (q < n_live) {
                const unsigned info = d0[(size_t)q * DESC_WORDS + D_INFO];
                if (info & 1u) {
                    go_mask |= 1u << q;
                    cnt[q] = (int)((info >> 16) & 255u);
                    total += (cnt[q] + 3) & ~3;
                }
            }
        }
        off[BG] = total;
        if (go_mask == 0u) {
            continue;   // (no frame of the group runs the synthesis: nothing to leave behind)
        }
#pragma unroll
        for (int q = 0; q < BG; ++q) {
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                ws.out[q][32 * c + lane] = 0.0f;
            }
        }

        int n_k2 = 0;   // interpolated harmonics of the group (profiling counter)
#pragma unroll 1
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            int q = 0;
#pragma unroll
            for (int t = 1; t < BG; ++t) {
                q += (off[t] <= k) ? 1 : 0;
            }
            int ofs_q = off[0], cnt_q = cnt[0];
#pragma unroll
            for (int t = 1; t < BG; ++t) {
                if (q == t) {
                    ofs_q = off[t];
                    cnt_q = cnt[t];
                }
            }
            const int pos = k - ofs_q;
            const bool used = pos < cnt_q;
            float g = 0.f, c = 0.f, s = 0.f, cd = 0.f, sd = 0.f, dw0 = 0.f;
            int kind = 0, l = 0;
            if (used) {
                const uint32_t* d = d0 + (size_t)q * DESC_WORDS;
                const unsigned id = reinterpret_cast<const unsigned char*>(d + D_KIND)[pos];
                kind = (int)(id & 3u);
                l = (int)(id >> 2);
                g = __uint_as_float(d[D_G + pos]);
                c = __uint_as_float(d[D_C + pos]);
                s = __uint_as_float(d[D_S + pos]);
                cd = __uint_as_float(d[D_CD + pos]);
                sd = __uint_as_float(d[D_SD + pos]);
                dw0 = __uint_as_float(d[D_DW0]);
            }
            const bool k2lane = (kind == 2);
            const unsigned k2mask = __ballot_sync(FULL, k2lane);
            n_k2 += __popc(k2mask);
            const float* Wb = vwin + ((kind == 0) ? WIN_PREV : 0);
            const float gg = k2lane ? 0.0f : g;   // interpolated slots are written by whoever renders them
            const float rec_c = c, rec_s = s;     // (the recurrence below rotates c and s on every lane)
            // phase B's map of the pass: which frame owns each four-slot group of the tile (a frame's slots start at a multiple
            // of four, so a group never straddles frames): bit 4g of (fq0, fq1) = frame of group g, fchg = groups where it changes
            static_assert(BG <= 8, "three bits per group");
            const unsigned fq0 = __ballot_sync(FULL, (q & 1) != 0), fq1 = __ballot_sync(FULL, (q & 2) != 0);
            const unsigned fq2 = (BG > 4) ? __ballot_sync(FULL, (q & 4) != 0) : 0u;
            const unsigned fchg = (fq0 ^ (fq0 << 4)) | (fq1 ^ (fq1 << 4)) | (fq2 ^ (fq2 << 4));
            const int ng = min(8, (total - base) >> 2);   // groups of this pass that hold slots
#pragma unroll 1
            for (int ch = 0; ch < 5; ++ch) {
                // phase A: 32 oscillator steps, sixteen per loop body (the body stays in the L0 instruction cache)
                const float4* W4 = reinterpret_cast<const float4*>(Wb) + 8 * ch;
                float* tp = tile + lane;
#pragma unroll 1
                for (int h = 0; h < 32 / BANK_STEPS; ++h) {
#pragma unroll
                    for (int n4 = 0; n4 < BANK_STEPS / 4; ++n4) {
                        const float4 w4 = W4[(BANK_STEPS / 4) * h + n4];
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (!k2lane) {
                                tp[(4 * n4 + i) * BT_STRIDE] = (gg * wv[i]) * c;   // row n = BANK_STEPS h + 4 n4 + i
                            }
                            const float cn = (c * cd) - (s * sd);
                            const float sn = (s * cd) + (c * sd);
                            c = cn;
                            s = sn;
                        }
                    }
                    tp += BANK_STEPS * BT_STRIDE;
                }
                // phase-interpolated harmonics of this pass: lane = sample (mbelib.c:953-968); the slot's lane holds its record
                if (k2mask) {
                    const int n = 32 * ch + lane;
                    const float fn = (float)n, fr = nfrac[n];
                    const int nn = n * n;
                    for (unsigned m = k2mask; m; m &= m - 1u) {
                        const int sl = __ffs(m) - 1;
                        const float a1 = __shfl_sync(FULL, g, sl), phi = __shfl_sync(FULL, rec_c, sl),
                                    pM = __shfl_sync(FULL, rec_s, sl), cM = __shfl_sync(FULL, cd, sl),
                                    dw = __shfl_sync(FULL, dw0, sl);
                        const int ll = __shfl_sync(FULL, l, sl);
                        const float th = phi + (a1 * fn) + ((dw * (float)(ll * nn)) / (float)(2 * NS));
                        const float am = pM + (fr * (cM - pM));
                        tile[lane * BT_STRIDE + sl] = 2.0f * am * dev_cosf(th);
                    }
                }
                __syncwarp();
                // phase B: lane = sample; the tile's groups in order, each added to the running sum of the frame that owns it
                // (list order = the reference's summation order); the sum moves to the next frame's row where the owner changes
                {
                    const float4* row = reinterpret_cast<const float4*>(tile + lane * BT_STRIDE);
                    float* op = &ws.out[(fq0 & 1u) | ((fq1 & 1u) << 1) | ((fq2 & 1u) << 2)][32 * ch + lane];
                    float a = *op;
#pragma unroll
                    for (int gi = 0; gi < 8; ++gi) {
                        if (gi < ng) {
                            if (gi > 0 && ((fchg >> (4 * gi)) & 1u)) {
                                *op = a;
                                op = &ws.out[((fq0 >> (4 * gi)) & 1u) | (((fq1 >> (4 * gi)) & 1u) << 1) |
                                             (((fq2 >> (4 * gi)) & 1u) << 2)][32 * ch + lane];
                                a = *op;
                            }
                            const float4 v = row[gi];
                            a += v.x;
                            a += v.y;
                            a += v.z;
                            a += v.w;
                        }
                    }
                    *op = a;
                }
                __syncwarp();
            }
        }
        if (A.counters && lane == 0) {
            int slots = 0;
#pragma unroll
            for (int q = 0; q < BG; ++q) {
                slots += cnt[q];
            }
            atomicAdd(&A.counters[0], (unsigned long long)slots);
            atomicAdd(&A.counters[1], (unsigned long long)n_k2);
            atomicAdd(&A.counters[2], (unsigned long long)__popc(go_mask));
        }
        // the frames' voiced samples replace their (now dead) slot records
#pragma unroll
        for (int q = 0; q < BG; ++q) {
            if ((go_mask >> q) & 1u) {
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    d0[(size_t)q * DESC_WORDS + D_VOICED + 32 * c + lane] = __float_as_uint(ws.out[q][32 * c + lane]);
                }
            }
        }
    }
}

// ---- unvoiced kernel ---------------------------------------------------------------------------------------------------
struct __align__(16) UnvWS {
    float a[324];
    float b[324];
    float scale[132];
};

__device__ __forceinline__ void uw_copy(uint32_t* dst, const uint32_t* src, int lane) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = src[UW_WORD + 32 * k + lane];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        dst[UW_WORD + 32 * k + lane] = v[k];
    }
}

// band scales + backward transform for one frame (unvoiced_shape on descriptor data)
__device__ __forceinline__ void split_unvoiced_shape(UnvWS& ws, const uint32_t* d, int L, float w0, unsigned cu_lo, unsigned cu_hi,
                                                     const BlockTables* bt, int lane) {
    float* A = ws.a;
    float* scale = ws.scale;
    const float mult = (256.0f / (2.0f * 3.14159265358979323846f)) * w0;
    for (int l = 1 + lane; l <= L; l += 32) {
        int a = (int)ceilf((l - 0.5f) * mult);
        int b = (int)ceilf((l + 0.5f) * mult);
        if (a < 0) {
            a = 0;
        }
        if (b > NFFT / 2) {
            b = NFFT / 2;
        }
        const bool uv = ((l <= 32 ? cu_lo >> (l - 1) : cu_hi >> (l - 33)) & 1u) != 0u;
        if (uv && b > a) {
            float num = 0.0f;
            int s = a;
            if (s == 0) {
                num += A[1] * A[1];
                s = 1;
            }
            for (int bin = s; bin < b; ++bin) {
                const Cpx v = ld2(A + 2 * bin);
                num += (v.re * v.re) + (v.im * v.im);
            }
            if (num > 1e-10f) {
                const float sc = 146.17696f * __uint_as_float(d[D_CML + l - 1]) / sqrtf(num / (float)(b - a));
                for (int bin = a; bin < b; ++bin) {
                    scale[bin] = sc;
                }
            }
        }
    }
    __syncwarp();
    rfft256_backward(A, ws.b, bt->tw, scale, lane);
}

// the frame's windowed noise buffer from the descriptor (make_noise's FFT input)
__device__ __forceinline__ void split_noise(UnvWS& ws, const uint32_t* d, const BlockTables* bt, int lane) {
    float* A = ws.a;
    const float seed = __uint_as_float(d[D_SEED]);
    if (seed < 0.0f) {
        for (int i = lane; i < NFFT; i += 32) {
            A[i] = 0.0f;
        }
        return;
    }
    constexpr unsigned A32 = lcg_pow_a(32), C32 = lcg_pow_c(32);
    unsigned sc = ((unsigned)seed) % 53125u;
    const uint2 jl = bt->uv_jump[lane];
    float ov[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ov[r] = __uint_as_float(d[D_OV + 32 * r + lane]);
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int i = 32 * c + lane;
        const unsigned st = (jl.x * sc + jl.y) % 53125u;
        sc = (A32 * sc + C32) % 53125u;
        A[96 + i] = (float)st * bt->uvwin[96 + i];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = 32 * r + lane;
        A[i] = ov[r] * bt->uvwin[i];
    }
}

// one warp = one stream, frames in order (the overlap-add needs the previous frame's block)
__global__ void __launch_bounds__(U_WARPS * 32, U_MINB) mbe_split_unvoiced_kernel(const SynthArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockTables* bt = reinterpret_cast<BlockTables*>(smem_raw);
    UnvWS* wsa = reinterpret_cast<UnvWS*>(smem_raw + sizeof(BlockTables));
    const DevTables* T = A.tab;
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            bt->tw[i] = T->tw[i];
            bt->uvwin[i] = T->uvwin[i];
        }
        for (int i = threadIdx.x; i < 160; i += blockDim.x) {
            bt->wola_wp[i] = T->wola_wp[i];
            bt->wola_wc[i] = T->wola_wc[i];
            bt->wola_den[i] = T->wola_den[i];
        }
        if (threadIdx.x < 32) {
            bt->uv_jump[threadIdx.x] = make_uint2(T->uvA[threadIdx.x], T->uvC[threadIdx.x]);
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    UnvWS& ws = wsa[warp];
    const int s = blockIdx.x * U_WARPS + warp;
    const bool live = s < A.n_streams;
    uint32_t* gs = A.state + (size_t)(A.first_stream + (live ? s : 0)) * STATE_WORDS;
    uint32_t* h_cur = gs;
    uint32_t* h_prev = gs + PARMS_WORDS;
    uint32_t* h_enh = gs + 2 * PARMS_WORDS;
    uint32_t* h_spill = gs + SPILL_WORD;
#pragma unroll 1
    for (int f = 0; f < A.n_frames; ++f) {
#if MBE_U_LOCKSTEP
        __syncthreads();
#endif
        if (!live) {
            continue;
        }
        const size_t idx = (size_t)(A.io_base + s) * A.n_frames + f;
        const uint32_t* d = A.desc + ((size_t)s * A.n_frames + f) * DESC_WORDS;
        unsigned ops = d[D_OPS];
#pragma unroll 1
        while (ops) {
            const unsigned op = ops & 15u;
            ops >>= 4;
            if (op == OP_SYNTH) {
                const unsigned info = d[D_INFO];
                const int L = (int)((info >> 8) & 255u);
                float tail[4], voiced[5];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tail[q] = __uint_as_float(h_enh[UW_WORD + 128 + 32 * q + lane]);
                }
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    voiced[q] = __uint_as_float(d[D_VOICED + 32 * q + lane]);
                }
                split_noise(ws, d, bt, lane);
                for (int i = lane; i < 129; i += 32) {
                    ws.scale[i] = 0.0f;
                }
                __syncwarp();
                rfft256_forward(ws.a, ws.b, bt->tw, lane);
                split_unvoiced_shape(ws, d, L, __uint_as_float(d[D_CW0]), d[D_CU], d[D_CU + 1], bt, lane);
                const float* Ab = ws.a;
                const float inv = 1.0f / (float)NFFT;
#pragma unroll
                for (int cc = 0; cc < 5; ++cc) {
                    const int n = 32 * cc + lane;
                    const float den = bt->wola_den[n];
                    const float ps = (cc < 4) ? tail[cc < 4 ? cc : 0] : 0.0f;
                    const float cs = (n - 32 >= 0) ? (Ab[n - 32] * inv) : 0.0f;
                    float v = voiced[cc];
                    if (den > 1e-10f) {
                        v += ((bt->wola_wp[n] * ps) + (bt->wola_wc[n] * cs)) / den;
                    }
                    if (v > MBE_CLIP_F) {
                        v = MBE_CLIP_F;
                    } else if (v < -MBE_CLIP_F) {
                        v = -MBE_CLIP_F;
                    }
                    const size_t o = idx * NS + n;
                    if (A.pcmf) {
                        A.pcmf[o] = ref_nan(v * A.pcmf_scale);
                    }
                    if (A.pcm) {
                        A.pcm[o] = float_to_short(v);
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t v = __float_as_uint(Ab[32 * q + lane] * inv);
                    h_cur[UW_WORD + 32 * q + lane] = v;
                    h_enh[UW_WORD + 32 * q + lane] = v;
                }
                __syncwarp();
            } else if (op == OP_PREV_FROM_CUR) {
                uw_copy(h_prev, h_cur, lane);
            } else if (op == OP_CUR_FROM_PREV) {
                uw_copy(h_cur, h_prev, lane);
            } else if (op == OP_ENH_FROM_CUR) {
                uw_copy(h_enh, h_cur, lane);
            } else if (op == OP_CUR_FROM_ENH) {
                uw_copy(h_cur, h_enh, lane);
            } else if (op == OP_SPILL_FROM_CUR) {
                uw_copy(h_spill, h_cur, lane);
            } else if (op == OP_CUR_FROM_SPILL) {
                uw_copy(h_cur, h_spill, lane);
            } else if (op == OP_ZERO_CUR) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    h_cur[UW_WORD + 32 * q + lane] = 0u;
                }
            }
        }
    }
}

}  // namespace mbe
