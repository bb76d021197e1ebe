// log10f_emul.h -- glibc's single-precision log10 restated in plain IEEE-754 operations, for host and device.
//
// The height-score bin of pingpongpro.cpp:591-594 goes through the HOST's log10f, so the bin thresholds depend on the
// libm the reference was linked against.  Computing them on the host costs a device -> host -> device
// round trip in the middle of every pass.  The device therefore evaluates the same function itself -- glibc 2.39:
//   log10f(x) = (y * log10_2lo + ivln10 * logf(m)) + y * log10_2hi,   x = m * 2^y, m in [1, 2)     (e_log10f.c)
//   logf      = the table-driven routine of ARM's optimized-routines (16 entries of {1/c, log c}, a cubic in double
//               precision, result rounded once to float)                                            (e_logf.c)
// Both the FMA and the non-FMA build of that logf (glibc selects one at run time) round to the same float for every
// argument in [1, 2^49) -- checked exhaustively (411 M floats, tests/test_host_tables.py runs a strided version) -- and
// so does this restatement.  The speculation is VERIFIED on every pass all the same: the host computes the table with
// its own libm while the device works on and compares (api.cu: settle); a mismatch (another libm) switches the context
// to host-computed tables for good.  Domain: finite x >= 1 (height-score products are >= 1).
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define PPP_HD __host__ __device__ inline
#else
#define PPP_HD inline
#endif

namespace pppk {

PPP_HD double emul_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
PPP_HD double emul_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
PPP_HD double emul_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
PPP_HD float emul_fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;  // (no contraction, no excess precision)
    return r;
#endif
}
PPP_HD float emul_fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
PPP_HD float emul_fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b;
    return r;
#endif
}
PPP_HD uint32_t emul_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
#endif
}
PPP_HD float emul_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// logf for a float in [1, 2)
PPP_HD float emul_logf_mantissa(float m) {
    // {1/c, log(c)} for the 16 subintervals of [0x1.66p-1, 0x1.66p0) (the published table of the routine)
    const double kInvC[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0, 0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,
                              0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0, 0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
                              0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
    const double kLogC[16] = {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3, -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,
                              -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4, -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
                              0x1.526e57720db08p-3,  0x1.bc2860d22477p-3,   0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};
    const double kLn2 = 0x1.62e42fefa39efp-1, kA0 = -0x1.00ea348b88334p-2, kA1 = 0x1.5575b0be00b6ap-2, kA2 = -0x1.ffffef20a4123p-2;
    const uint32_t ix = emul_bits(m);
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)((int32_t)tmp >> 23);
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double z = (double)emul_float(iz);
    const double r = emul_fma(z, kInvC[i], -1.0);
    const double y0 = emul_fma((double)k, kLn2, kLogC[i]);
    double y = emul_fma(kA1, r, kA2);
    const double r2 = emul_dmul(r, r);
    const double t = emul_dadd(r, y0);
    y = emul_fma(kA0, r2, y);
    return (float)emul_fma(r2, y, t);
}

// log10f for finite x >= 1
PPP_HD float emul_log10f(float x) {
    const uint32_t hx = emul_bits(x);
    const int k = (int)(hx >> 23) - 127;
    const float y = (float)k;
    const float m = emul_float((hx & 0x007fffffu) | 0x3f800000u);
    const float z = emul_fadd(emul_fmul(4.3429449201e-01f, emul_logf_mantissa(m)), emul_fmul(y, 7.9034151668e-07f));  // ivln10 * logf(m) + y * log10_2lo
    return emul_fadd(z, emul_fmul(y, 3.0102920532e-01f));                                                              // + y * log10_2hi
}

// heightScoreBin of pingpongpro.cpp:591-594 for one float product (float log10 -> float divide -> float multiply by
// 999 -> double add 0.5 -> truncate), with the emulated log10f.
PPP_HD int emul_height_score_bin(float hs, float maxHeightScore) {
    const float t = emul_fmul(emul_fdiv(emul_log10f(hs), maxHeightScore), 999.0f);
    return (int)emul_dadd(0.5, (double)t);
}

// thresholds[b-1] (b = 1..999): smallest float hs in [1, maxFreq^2] with emul_height_score_bin(hs) >= b, +inf when
// maxFreq^2 does not reach bin b -- the table of host_tables.h: computeThresholds, by bisection on float bit patterns.
PPP_HD float emul_bin_threshold(int b, float maxFreq) {
    const float top = emul_fmul(maxFreq, maxFreq);
    const float maxHs = emul_log10f(top);  // :551
    if (emul_height_score_bin(top, maxHs) < b) return emul_float(0x7f800000u);
    uint32_t lo = 0x3f800000u, hi = emul_bits(top);  // invariant: bin(hi) >= b
    {   // a narrow bracket around 10^(maxHs * (b - 0.5) / 999), kept only if it really brackets the step (it only saves probes)
        const double est = pow(10.0, (double)maxHs * ((double)b - 0.5) / 999.0);
        const float fl = (float)(est * (1.0 - 3e-5)), fh = (float)(est * (1.0 + 3e-5));
        if (fl > 1.0f && fh < top) {
            const uint32_t bl = emul_bits(fl), bh = emul_bits(fh);
            if (emul_height_score_bin(fl, maxHs) < b && emul_height_score_bin(fh, maxHs) >= b) { lo = bl + 1u; hi = bh; }
        }
    }
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (emul_height_score_bin(emul_float(mid), maxHs) >= b) hi = mid; else lo = mid + 1u;
    }
    return emul_float(hi);
}

}  // namespace pppk
