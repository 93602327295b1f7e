// hk_detmath.h — one deterministic f32 libm, compiled into BOTH libhikari_cuda.so (device code) and the CPU oracle.
//
// Why: several stages of the path seed private RNGs from the BITS of floats (delta tracking delta-tracking.jl:28-45, ratio
// tracking intersection.jl:455, the LayeredBxDF walk spectral-eval.jl:1318, MixMaterial mix-material.jl:114-158).  A 1-ulp
// difference between two libms (glibc vs CUDA vs Julia's own) upstream of such a hash reseeds the walk, so two
// implementations can then only agree in distribution.  Every transcendental on the path therefore goes through the
// functions below, which use nothing but IEEE-754 operations that are correctly rounded on both sides: + - * / fma,
// floor / rint, float <-> int conversion of in-range values and bit casts.  Same source, same operation order, same bits on
// x86-64 (g++ -ffp-contract=off; explicit fma -> vfmadd or glibc's exact fmaf) and on sm_100a (nvcc -fmad=false; explicit
// __fmaf_rn), denormals included (neither side flushes).
//
// Accuracy (tests/test_detmath.py, against double-precision libm): expf / logf <= 1.5 ulp, sinf / cosf <= 2 ulp (2.5 up to 3e4) on the ranges the
// path uses, coshf / atanhf <= 3 ulp, powf <= 1 ulp (evaluated in double).  sin / cos reduce with a three-constant
// Cody-Waite scheme up to |x| = 2^15 and in double above that (accurate to ~|x| * 2^-52: deterministic for every input, not
// accurate for astronomically large ones -- the path's arguments are 2 pi u and pi/4 ratios).
// Algorithms: the classic fdlibm / Cephes kernels restated (polynomial coefficients are the published ones).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define DM_FN __host__ __device__ __forceinline__
// HK_NOINLINE_LIBM: exp / log / sin / cos as real functions in device code (the kernels that use them are several times the 32 KB
// instruction cache; measured on B200: C4 +3.3 %, C2 / C3 / C5 +0.3-0.9 %)
#ifndef HK_NOINLINE_LIBM
#define HK_NOINLINE_LIBM 1
#endif
#if HK_NOINLINE_LIBM
#define DM_FN_BIG static __host__ __device__ __noinline__
#else
#define DM_FN_BIG DM_FN
#endif
#else
#include <string.h>
#define DM_FN static inline
#define DM_FN_BIG DM_FN
#endif

#if defined(__CUDA_ARCH__)
DM_FN float dm_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
DM_FN double dm_fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
DM_FN uint32_t dm_bits(float x) { return __float_as_uint(x); }
DM_FN float dm_float(uint32_t u) { return __uint_as_float(u); }
DM_FN uint64_t dm_bitsd(double x) { return (uint64_t)__double_as_longlong(x); }
DM_FN double dm_double(uint64_t u) { return __longlong_as_double((long long)u); }
DM_FN float dm_floor(float x) { return floorf(x); }
DM_FN float dm_rint(float x) { return rintf(x); }
DM_FN double dm_rintd(double x) { return rint(x); }
#else
DM_FN float dm_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
DM_FN double dm_fmad(double a, double b, double c) { return __builtin_fma(a, b, c); }
DM_FN uint32_t dm_bits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
DM_FN float dm_float(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
DM_FN uint64_t dm_bitsd(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
DM_FN double dm_double(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
DM_FN float dm_floor(float x) { return __builtin_floorf(x); }
DM_FN float dm_rint(float x) { return __builtin_rintf(x); }          // round-to-nearest-even (default rounding mode)
DM_FN double dm_rintd(double x) { return __builtin_rint(x); }
#endif

#define DM_INF_BITS 0x7f800000u
#define DM_NAN_BITS 0x7fc00000u

// ---- expf: x = k ln2 + r, |r| <= ln2/2; e^r by the Cephes degree-7 polynomial; 2^k applied in two exact steps so that
// results in the denormal range round once ----------------------------------------------------------------------------
DM_FN_BIG float dm_expf(float x) {
    if (x != x) return x;
    if (x > 88.72284f) return dm_float(DM_INF_BITS);
    if (x < -104.0f) return 0.0f;
    const float kf = dm_floor(dm_fma(x, 1.44269504088896341f, 0.5f));
    float r = dm_fma(kf, -0.693359375f, x);              // ln2 split: 0.693359375 - 2.12194440e-4
    r = dm_fma(kf, 2.12194440e-4f, r);
    const float z = r * r;
    float p = 1.9875691500e-4f;
    p = dm_fma(p, r, 1.3981999507e-3f);
    p = dm_fma(p, r, 8.3334519073e-3f);
    p = dm_fma(p, r, 4.1665795894e-2f);
    p = dm_fma(p, r, 1.6666665459e-1f);
    p = dm_fma(p, r, 5.0000001201e-1f);
    p = dm_fma(p, z, r) + 1.0f;
    const int k = (int)kf;                                // |k| <= 151
    const int k1 = k / 2, k2 = k - k1;
    return (p * dm_float((uint32_t)(k1 + 127) << 23)) * dm_float((uint32_t)(k2 + 127) << 23);
}

// ---- logf: fdlibm e_logf.c ----------------------------------------------------------------------------------------
DM_FN_BIG float dm_logf(float x) {
    uint32_t ix = dm_bits(x);
    int k = 0;
    if (x != x) return x;
    if (ix >= 0x80000000u) return (ix << 1) == 0u ? dm_float(0xff800000u) : dm_float(DM_NAN_BITS);     // -0 -> -inf, x < 0 -> NaN
    if (ix == 0u) return dm_float(0xff800000u);
    if (ix == DM_INF_BITS) return x;
    if (ix < 0x00800000u) { x = x * 33554432.0f; ix = dm_bits(x); k = -25; }      // denormal: scale by 2^25 (exact)
    k += (int)(ix >> 23) - 127;
    ix &= 0x007fffffu;
    const uint32_t i = (ix + (0x95f64u << 3)) & 0x800000u;                         // mantissa >= sqrt(2): halve it
    const float m = dm_float(ix | (i ^ 0x3f800000u));
    k += (int)(i >> 23);
    const float f = m - 1.0f;
    const float dk = (float)k;
    const float s = f / (2.0f + f);
    const float z = s * s, w = z * z;
    const float t1 = w * dm_fma(w, 0.24279078841f, 0.40000972152f);
    const float t2 = z * dm_fma(w, 0.28498786688f, 0.66666662693f);
    const float R = t2 + t1;
    const float hfsq = 0.5f * f * f;
    return dm_fma(dk, 6.9313812256e-01f, -((hfsq - dm_fma(s, hfsq + R, dk * 9.0580006145e-06f)) - f));
}
// log(1 + y) with the classic correction term (u = 1 + y rounded; log(u) * y / (u - 1))
DM_FN float dm_log1pf(float y) {
    const float u = 1.0f + y;
    if (u == 1.0f) return y;
    if (!(u > 0.0f) || u != u || dm_bits(u) == DM_INF_BITS) return dm_logf(u);
    return dm_logf(u) * (y / (u - 1.0f));
}

// ---- sinf / cosf: n = rint(x 2/pi), r = x - n pi/2 (Cody-Waite, three constants, fused), fdlibm k_sinf / k_cosf
// polynomials on [-pi/4, pi/4] -------------------------------------------------------------------------------------------
DM_FN float dm_ksin(float r) {
    const float z = r * r;
    float p = 2.7557314297e-06f;
    p = dm_fma(p, z, -1.9841270114e-04f);
    p = dm_fma(p, z, 8.3333337680e-03f);
    p = dm_fma(p, z, -1.6666667163e-01f);
    return dm_fma(r * z, p, r);
}
DM_FN float dm_kcos(float r) {
    const float z = r * r;
    float p = -2.7557314297e-07f;
    p = dm_fma(p, z, 2.4801587642e-05f);
    p = dm_fma(p, z, -1.3888889225e-03f);
    p = dm_fma(p, z, 4.1666667908e-02f);
    return dm_fma(z * z, p, dm_fma(z, -0.5f, 1.0f));
}
// quadrant n (mod 4) and reduced argument
DM_FN int dm_rem_pio2(float x, float& r) {
    const float ax = x < 0.0f ? -x : x;
    if (ax <= 32768.0f) {
        const float nf = dm_rint(x * 0.63661977236758134f);
        r = dm_fma(nf, -1.5707855225e+00f, x);
        r = dm_fma(nf, -1.0804273188e-05f, r);
        r = dm_fma(nf, -6.0770999344e-11f, r);
        return (int)nf & 3;
    }
    // large arguments: the same reduction in double, repeated while the remainder is still out of range (|x| > ~2^50 is
    // reduced in up to three rounds: deterministic and bounded, not accurate -- see the header)
    double rd = (double)x, qsum = 0.0;
    for (int it = 0; it < 4; it++) {
        const double nd = dm_rintd(rd * 0.63661977236758134308);
        rd = dm_fmad(nd, -1.57079632673412561417e+00, rd);
        rd = dm_fmad(nd, -6.07710050650619224932e-11, rd);
        qsum += nd - 4.0 * dm_rintd(nd * 0.25);             // quadrant mod 4, exact (nd is an integer-valued double)
        if (rd <= 0.7853981633974484 && rd >= -0.7853981633974484) break;
    }
    if (!(rd <= 0.7853981633974484 && rd >= -0.7853981633974484)) rd = 0.0;
    r = (float)rd;
    return (int)qsum & 3;
}
DM_FN_BIG float dm_sinf(float x) {
    if (x != x || dm_bits(x < 0.0f ? -x : x) == DM_INF_BITS) return dm_float(DM_NAN_BITS);
    float r; const int n = dm_rem_pio2(x, r);
    const float v = (n & 1) ? dm_kcos(r) : dm_ksin(r);
    return (n & 2) ? -v : v;
}
DM_FN_BIG float dm_cosf(float x) {
    if (x != x || dm_bits(x < 0.0f ? -x : x) == DM_INF_BITS) return dm_float(DM_NAN_BITS);
    float r; const int n = dm_rem_pio2(x, r);
    const float v = (n & 1) ? dm_ksin(r) : dm_kcos(r);
    return ((n + 1) & 2) ? -v : v;
}

// ---- coshf / atanhf ------------------------------------------------------------------------------------------------------
DM_FN float dm_coshf(float x) {
    if (x != x) return x;
    const float ax = x < 0.0f ? -x : x;
    if (ax > 88.0f) { const float e = dm_expf(ax - 44.0f); return (0.5f * e) * 1.2851600114359308e19f; }   // e^44, avoids early overflow
    const float e = dm_expf(ax);
    return 0.5f * e + 0.5f / e;
}
DM_FN float dm_atanhf(float x) {
    if (x != x) return x;
    const float ax = x < 0.0f ? -x : x;
    if (ax > 1.0f) return dm_float(DM_NAN_BITS);
    if (ax == 1.0f) return x < 0.0f ? dm_float(0xff800000u) : dm_float(DM_INF_BITS);
    const float h = 0.5f * dm_log1pf((2.0f * ax) / (1.0f - ax));
    return x < 0.0f ? -h : h;
}

// ---- powf, evaluated in double (log and exp to ~1e-13 relative), rounded once to f32 ---------------------------------------
DM_FN double dm_log_d(double x) {       // x > 0, finite, normal (callers pass f32 values widened to double)
    const uint64_t ix = dm_bitsd(x);
    int k = (int)(ix >> 52) - 1023;
    uint64_t mant = ix & 0x000fffffffffffffull;
    double m = dm_double(mant | 0x3ff0000000000000ull);
    if (m > 1.4142135623730951) { m = m * 0.5; k += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 1.0 / 19.0;
    p = dm_fmad(p, z, 1.0 / 17.0);
    p = dm_fmad(p, z, 1.0 / 15.0);
    p = dm_fmad(p, z, 1.0 / 13.0);
    p = dm_fmad(p, z, 1.0 / 11.0);
    p = dm_fmad(p, z, 1.0 / 9.0);
    p = dm_fmad(p, z, 1.0 / 7.0);
    p = dm_fmad(p, z, 1.0 / 5.0);
    p = dm_fmad(p, z, 1.0 / 3.0);
    p = dm_fmad(p, z, 1.0);
    return dm_fmad((double)k, 0.69314718055994530942, 2.0 * s * p);
}
DM_FN double dm_exp_d(double t) {       // |t| < 745
    const double kf = dm_rintd(t * 1.44269504088896340736);
    double r = dm_fmad(kf, -6.93147180369123816490e-01, t);
    r = dm_fmad(kf, -1.90821492927058770002e-10, r);
    double p = 1.0 / 479001600.0;
    p = dm_fmad(p, r, 1.0 / 39916800.0);
    p = dm_fmad(p, r, 1.0 / 3628800.0);
    p = dm_fmad(p, r, 1.0 / 362880.0);
    p = dm_fmad(p, r, 1.0 / 40320.0);
    p = dm_fmad(p, r, 1.0 / 5040.0);
    p = dm_fmad(p, r, 1.0 / 720.0);
    p = dm_fmad(p, r, 1.0 / 120.0);
    p = dm_fmad(p, r, 1.0 / 24.0);
    p = dm_fmad(p, r, 1.0 / 6.0);
    p = dm_fmad(p, r, 0.5);
    p = dm_fmad(p, r, 1.0);
    p = dm_fmad(p, r, 1.0);
    const int k = (int)kf;                  // |k| <= 1075
    const int k1 = k / 2, k2 = k - k1;
    return (p * dm_double((uint64_t)(k1 + 1023) << 52)) * dm_double((uint64_t)(k2 + 1023) << 52);
}
DM_FN float dm_powf(float x, float y) {
    if (y == 0.0f || x == 1.0f) return 1.0f;
    if (x != x || y != y) return dm_float(DM_NAN_BITS);
    const float ay = y < 0.0f ? -y : y;
    const bool y_int = ay >= 8388608.0f || dm_floor(ay) == ay;
    const bool y_odd = y_int && ay < 16777216.0f && ((int)ay & 1);
    const float ax = x < 0.0f ? -x : x;
    const bool neg = (dm_bits(x) >> 31) != 0u;
    if (ax == 0.0f) { const float v = y < 0.0f ? dm_float(DM_INF_BITS) : 0.0f; return (neg && y_odd) ? -v : v; }
    if (neg && !y_int) return dm_float(DM_NAN_BITS);
    float mag;
    if (dm_bits(ay) == DM_INF_BITS) mag = (ax == 1.0f) ? 1.0f : (((ax > 1.0f) == (y > 0.0f)) ? dm_float(DM_INF_BITS) : 0.0f);
    else if (dm_bits(ax) == DM_INF_BITS) mag = y > 0.0f ? dm_float(DM_INF_BITS) : 0.0f;
    else {
        const double t = (double)y * dm_log_d((double)ax);      // f32 denormals are normal doubles
        mag = t > 89.0 ? dm_float(DM_INF_BITS) : (t < -104.0 ? 0.0f : (float)dm_exp_d(t));
    }
    return (neg && y_odd) ? -mag : mag;
}
