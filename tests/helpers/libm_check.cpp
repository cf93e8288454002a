// Host-side check of the device libm ports (mbelib-neo_b200/csrc/mbe_libm.cuh) against the host glibc.
// Build: g++ -O2 -ffp-contract=off -fopenmp -I mbelib-neo_b200/csrc tests/helpers/libm_check.cpp -o libm_check
// Usage: libm_check <stride>   (stride 1 = every one of the 2^32 float bit patterns)
// Prints one line per function: name, values tested, mismatches (bit-level, NaNs compared as a class).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "mbe_libm.cuh"
#include "mbe_exp2_tab.inc"

static inline bool same(float a, float b) {
    if (std::isnan(a) && std::isnan(b)) return true;
    return mbelibm::f2u(a) == mbelibm::f2u(b);
}

int main(int argc, char** argv) {
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    uint64_t bad_sin = 0, bad_cos = 0, bad_sc = 0, bad_e2 = 0, bad_e = 0, n = 0;
#pragma omp parallel for reduction(+ : bad_sin, bad_cos, bad_sc, bad_e2, bad_e, n) schedule(static)
    for (uint64_t u = 0; u < (1ull << 32); u += stride) {
        float x = mbelibm::u2f((uint32_t)u);
        float s, c, rs, rc;
        mbelibm::sincosf_glibc(x, &s, &c);
        sincosf(x, &rs, &rc);
        if (!same(s, rs) || !same(c, rc)) bad_sc++;
        if (!same(mbelibm::sinf_glibc(x), sinf(x))) bad_sin++;
        if (!same(mbelibm::cosf_glibc(x), cosf(x))) bad_cos++;
        if (!same(mbelibm::exp2f_glibc(x, mbe_exp2_tab), exp2f(x))) bad_e2++;
        if (!same(mbelibm::expf_glibc(x, mbe_exp2_tab), expf(x))) bad_e++;
        n++;
    }
    printf("tested %llu\nsincosf %llu\nsinf %llu\ncosf %llu\nexp2f %llu\nexpf %llu\n", (unsigned long long)n,
           (unsigned long long)bad_sc, (unsigned long long)bad_sin, (unsigned long long)bad_cos,
           (unsigned long long)bad_e2, (unsigned long long)bad_e);
    return (bad_sc | bad_sin | bad_cos | bad_e2 | bad_e) ? 1 : 0;
}
