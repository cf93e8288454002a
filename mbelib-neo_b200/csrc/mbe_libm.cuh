// Device-side single-precision transcendental functions that reproduce glibc 2.39's results.
//
// Why: the reference (arancormonk/mbelib-neo) calls libm's sincosf/cosf/sinf/exp2f/expf/log2f
// (call sites: src/core/mbelib.c:415,704,965,978-1011; src/imbe/imbe7200x4400.c:202,244,348;
// src/ambe/ambe3600x2450.c:440-449; src/ambe/ambe3600x2400.c:238,478-487;
// src/core/mbe_adaptive.c:186).  The PCM parity bar (>= 99.99 % of int16 samples exact) cannot be
// met with CUDA's 1-2 ulp intrinsics, so these routines restate the PUBLISHED algorithms glibc uses
// for the flt-32 functions (the ARM "optimized-routines" designs: double-precision polynomial
// kernels, one final rounding to float).  glibc is an external dependency of the reference (not in
// /root/reference); pinned version: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5).  tests/test_libm_port.py
// checks every function here against the host libm on dense/exhaustive argument sets.
//
// x86-64 glibc dispatches these functions to builds of the same C code compiled with -mfma, where the
// compiler fuses every multiply feeding an add.  The ports therefore spell the fused operations out
// with MBE_FMA (one rounding) exactly where that build has them; everything else is a plain
// multiply or add (the translation unit is compiled with -fmad=false, so nothing else is contracted).
// With this placement tests/helpers/libm_check.cpp finds 0 mismatches over all 2^32 arguments for
// each function on an FMA-capable host (and <= 34 of 2^32 against the non-FMA glibc variants).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MBE_HD __host__ __device__ __forceinline__
#else
#define MBE_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define MBE_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define MBE_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

namespace mbelibm {

MBE_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } v; v.f = f; return v.u;
#endif
}
MBE_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } v; v.u = u; return v.f;
#endif
}
MBE_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } v; v.d = d; return v.u;
#endif
}
MBE_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } v; v.u = u; return v.d;
#endif
}

// ---------------------------------------------------------------------------------------------
// sin/cos: quadrant reduction by pi/2, degree-8 cosine / degree-7 sine polynomials in double.
// ---------------------------------------------------------------------------------------------
struct SinCosCoef {
    double c0, c1, c2, c3, c4;  // cosine polynomial
    double s1, s2, s3;          // sine polynomial
};

// quadrant-pair 0 (n&2 == 0) and 1 (cosine coefficients negated)
#define MBE_SC_C0 0x1p0
#define MBE_SC_C1 -0x1.ffffffd0c621cp-2
#define MBE_SC_C2 0x1.55553e1068f19p-5
#define MBE_SC_C3 -0x1.6c087e89a359dp-10
#define MBE_SC_C4 0x1.99343027bf8c3p-16
#define MBE_SC_S1 -0x1.555545995a603p-3
#define MBE_SC_S2 0x1.1107605230bc4p-7
#define MBE_SC_S3 -0x1.994eb3774cf24p-13
#define MBE_SC_HPI_INV 0x1.45F306DC9C883p+23  /* 2/pi * 2^24 */
#define MBE_SC_HPI 0x1.921FB54442D18p0        /* pi/2 */
#define MBE_SC_PI63 0x1.921FB54442D18p-62     /* pi / 2^63 */

MBE_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ffu; }

// 2/pi as overlapping 32-bit windows of its binary expansion (byte stride), for |x| >= 120:
// window i covers bytes [i-3, i] of 0.A2F9836E 4E441529 FC2757D1 F534DDC0 DB629599 3C439041 ...
#if defined(__CUDACC__)
static __device__ const uint32_t k_inv_pio4_dev[24] = {
    0xa2u,       0xa2f9u,     0xa2f983u,   0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
    0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};
#endif

MBE_HD uint32_t inv_pio4_word(int i) {
#if defined(__CUDA_ARCH__)
    return k_inv_pio4_dev[i];
#else
    const uint64_t h0 = 0xA2F9836E4E441529ull, h1 = 0xFC2757D1F534DDC0ull, h2 = 0xDB6295993C439041ull;
    uint32_t w = 0;
    for (int b = i - 3; b <= i; ++b) {
        uint32_t byte = 0;
        if (b >= 0) {
            uint64_t q = (b < 8) ? h0 : (b < 16 ? h1 : h2);
            byte = (uint32_t)((q >> (8 * (7 - (b & 7)))) & 0xffu);
        }
        w = (w << 8) | byte;
    }
    return w;
#endif
}

MBE_HD double reduce_fast(double x, int* np) {
    double r = x * MBE_SC_HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return MBE_FMA(-(double)n, MBE_SC_HPI, x);
}

MBE_HD double reduce_large(uint32_t xi, int* np) {
    const int base = (int)((xi >> 26) & 15u);
    const int shift = (int)((xi >> 23) & 7u);
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (uint64_t)(uint32_t)(xi * inv_pio4_word(base + 0));
    res1 = (uint64_t)xi * inv_pio4_word(base + 4);
    res2 = (uint64_t)xi * inv_pio4_word(base + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * MBE_SC_PI63;
}

// polynomial evaluation; `neg` selects the table with negated cosine coefficients.
MBE_HD void sincos_poly(double x, double x2, int neg, int n, float* sinp, float* cosp) {
    const double sg = neg ? -1.0 : 1.0;
    const double c0 = sg * MBE_SC_C0, c1 = sg * MBE_SC_C1, c2 = sg * MBE_SC_C2, c3 = sg * MBE_SC_C3,
                 c4 = sg * MBE_SC_C4;
    double x3, x4, x5, x6, s, c, cc1, cc2, ss1;
    x4 = x2 * x2;
    x3 = x2 * x;
    cc2 = MBE_FMA(x2, c4, c3);
    ss1 = MBE_FMA(x2, MBE_SC_S3, MBE_SC_S2);
    cc1 = MBE_FMA(x2, c1, c0);
    x5 = x3 * x2;
    x6 = x4 * x2;
    s = MBE_FMA(x3, MBE_SC_S1, x);
    c = MBE_FMA(x4, c2, cc1);
    float sv = (float)MBE_FMA(x5, ss1, s);
    float cv = (float)MBE_FMA(x6, cc2, c);
    if (n & 1) {
        *cosp = sv;
        *sinp = cv;
    } else {
        *sinp = sv;
        *cosp = cv;
    }
}

MBE_HD float sin_poly(double x, double x2, int neg, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = MBE_FMA(x2, MBE_SC_S3, MBE_SC_S2);
        double x7 = x3 * x2;
        double s = MBE_FMA(x3, MBE_SC_S1, x);
        return (float)MBE_FMA(x7, s1, s);
    } else {
        const double sg = neg ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = MBE_FMA(x2, sg * MBE_SC_C4, sg * MBE_SC_C3);
        double c1 = MBE_FMA(x2, sg * MBE_SC_C1, sg * MBE_SC_C0);
        double x6 = x4 * x2;
        double c = MBE_FMA(x4, sg * MBE_SC_C2, c1);
        return (float)MBE_FMA(x6, c2, c);
    }
}

MBE_HD double quadrant_sign(int q) {
    // sign of sine in quadrants 0..3: +, -, -, +   (glibc table: {1,-1,-1,1})
    q &= 3;
    return (q == 1 || q == 2) ? -1.0 : 1.0;
}

MBE_HD void sincosf_glibc(float y, float* sinp, float* cosp) {
    double x = (double)y;
    int n;
    const uint32_t top = abstop12(y);
    if (top < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (top < abstop12(0x1p-12f)) {
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
        sincos_poly(x, x2, 0, 0, sinp, cosp);
    } else if (top < abstop12(120.0f)) {
        x = reduce_fast(x, &n);
        double s = quadrant_sign(n);
        sincos_poly(x * s, x * x, (n & 2) != 0, n, sinp, cosp);
    } else if (top < abstop12(u2f(0x7f800000u))) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        double s = quadrant_sign(n + sign);
        sincos_poly(x * s, x * x, ((n + sign) & 2) != 0, n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y;
    }
}

MBE_HD float sinf_glibc(float y) {
    double x = (double)y;
    int n;
    const uint32_t top = abstop12(y);
    if (top < abstop12(0x1.921FB6p-1f)) {
        double s = x * x;
        if (top < abstop12(0x1p-12f)) {
            return y;
        }
        return sin_poly(x, s, 0, 0);
    } else if (top < abstop12(120.0f)) {
        x = reduce_fast(x, &n);
        double s = quadrant_sign(n);
        return sin_poly(x * s, x * x, (n & 2) != 0, n);
    } else if (top < abstop12(u2f(0x7f800000u))) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        double s = quadrant_sign(n + sign);
        return sin_poly(x * s, x * x, ((n + sign) & 2) != 0, n);
    }
    return y - y;
}

MBE_HD float cosf_glibc(float y) {
    double x = (double)y;
    int n;
    const uint32_t top = abstop12(y);
    if (top < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (top < abstop12(0x1p-12f)) {
            return 1.0f;
        }
        return sin_poly(x, x2, 0, 1);
    } else if (top < abstop12(120.0f)) {
        x = reduce_fast(x, &n);
        double s = quadrant_sign(n);
        return sin_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
    } else if (top < abstop12(u2f(0x7f800000u))) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        double s = quadrant_sign(n + sign);
        return sin_poly(x * s, x * x, ((n + sign) & 2) != 0, n ^ 1);
    }
    return y - y;
}

// ---------------------------------------------------------------------------------------------
// exp2f / expf: 32-entry table of 2^(i/32), cubic in double.  `tab` must hold
// bits(2^(i/32)) - (i << 47) for i = 0..31 (see mbe_exp2_tab below).
// ---------------------------------------------------------------------------------------------
#define MBE_EXP2_C0 0x1.c6af84b912394p-5
#define MBE_EXP2_C1 0x1.ebfce50fac4f3p-3
#define MBE_EXP2_C2 0x1.62e42ff0c52d6p-1
#define MBE_EXP2_SHIFT_SCALED (0x1.8p+52 / 32.0)
#define MBE_EXP_SHIFT 0x1.8p+52
#define MBE_EXP_INVLN2N (0x1.71547652b82fep+0 * 32.0)

MBE_HD float exp2f_glibc(float x, const uint64_t* tab) {
    double xd = (double)x;
    uint32_t abstop = (f2u(x) >> 20) & 0x7ffu;
    if (abstop >= (f2u(128.0f) >> 20)) {
        if (f2u(x) == 0xff800000u) return 0.0f;
        if (abstop >= (0x7f800000u >> 20)) return x + x;
        if (x > 0.0f) return u2f(0x7f800000u);  // overflow
        if (x <= -150.0f) return 0.0f;          // underflow
    }
    double kd = xd + MBE_EXP2_SHIFT_SCALED;
    uint64_t ki = d2u(kd);
    kd -= MBE_EXP2_SHIFT_SCALED;
    double r = xd - kd;
    uint64_t t = tab[ki & 31u];
    t += ki << (52 - 5);
    double s = u2d(t);
    double z = MBE_FMA(MBE_EXP2_C0, r, MBE_EXP2_C1);
    double r2 = r * r;
    double y = MBE_FMA(MBE_EXP2_C2, r, 1.0);
    y = MBE_FMA(z, r2, y);
    y = y * s;
    return (float)y;
}

MBE_HD float expf_glibc(float x, const uint64_t* tab) {
    double xd = (double)x;
    uint32_t abstop = (f2u(x) >> 20) & 0x7ffu;
    if (abstop >= (f2u(88.0f) >> 20)) {
        if (f2u(x) == 0xff800000u) return 0.0f;
        if (abstop >= (0x7f800000u >> 20)) return x + x;
        if (x > 0x1.62e42ep6f) return u2f(0x7f800000u);
        if (x < -0x1.9fe368p6f) return 0.0f;
    }
    double kd = MBE_FMA(MBE_EXP_INVLN2N, xd, MBE_EXP_SHIFT);
    uint64_t ki = d2u(kd);
    kd -= MBE_EXP_SHIFT;
    double r = MBE_FMA(MBE_EXP_INVLN2N, xd, -kd);
    double z;
    uint64_t t = tab[ki & 31u];
    t += ki << (52 - 5);
    double s = u2d(t);
    z = MBE_FMA(MBE_EXP2_C0 / 32.0 / 32.0 / 32.0, r, MBE_EXP2_C1 / 32.0 / 32.0);
    double r2 = r * r;
    double y = MBE_FMA(MBE_EXP2_C2 / 32.0, r, 1.0);
    y = MBE_FMA(z, r2, y);
    y = y * s;
    return (float)y;
}

}  // namespace mbelibm
