// hk_math.cuh — float3 / 4-wide spectrum helpers, hashes, RNGs, ZSobol sampler (device side).
// The library is compiled with -fmad=false: every a*b+c below is two IEEE roundings, so results match
// the CPU restatement (-ffp-contract=off) and Julia's non-contracted CPU code up to libm differences.
// Reference lines: src/materials/spectral-eval.jl:575-815 (hashes, PCG32), src/sampler/sobol.jl (ZSobol),
// src/spectral/spectral.jl (SampledSpectrum{4}, wavelengths).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hk_detmath.h"

#define HK_DEV __device__ __forceinline__
// Code size matters on this path: fully inlined, the shading kernels were 110-670 KB of SASS each against a 32 KB L1.5
// instruction cache (ncu: stall_no_instruction up to 11 warps per issue).  Two remedies, both measured on B200:
// the light half of shading is its own kernel for large light sets (k_hit_lights), and selected helper groups can be compiled
// as real functions (HK_NI_*; 1 = __noinline__).  Same arithmetic in the same order either way: same bits.
#ifndef HK_NOINLINE_LIGHTS
#define HK_NOINLINE_LIGHTS 0       // light-BVH importance / descent, sample_light, env-map sampling
#endif
#ifndef HK_NOINLINE_BSDF
#define HK_NOINLINE_BSDF 1         // Trowbridge-Reitz D / Lambda / sample_wm, complex Fresnel x4 (measured: C5 shading 18.3 -> 14.9 ms, C2 1.52 -> 1.45; the L1.5 instruction cache is 32 KB)
#endif
#ifndef HK_NOINLINE_LAYERED
#define HK_NOINLINE_LAYERED 1      // coat_sample / coat_eval / coat_pdf / hg_sample_layer of the LayeredBxDF random walk
#endif
#ifndef HK_NOINLINE_SOBOL
#define HK_NOINLINE_SOBOL 0        // zsobol_1d / zsobol_2d
#endif
#ifndef HK_NOINLINE_DIGITS
#define HK_NOINLINE_DIGITS 0       // zsobol_digits (the permuted base-4 digit walk)
#endif
#ifndef HK_NOINLINE_MISC
#define HK_NOINLINE_MISC 1         // 4-wide exp, the NanoVDB root->leaf walk
#endif
#define HK_NI_ON static __device__ __noinline__
#define HK_NI_OFF __device__ __forceinline__
#if HK_NOINLINE_LIGHTS
#define HK_NI_LIGHTS HK_NI_ON
#else
#define HK_NI_LIGHTS HK_NI_OFF
#endif
#if HK_NOINLINE_BSDF
#define HK_NI_BSDF HK_NI_ON
#else
#define HK_NI_BSDF HK_NI_OFF
#endif
#if HK_NOINLINE_LAYERED
#define HK_NI_LAYERED HK_NI_ON
#else
#define HK_NI_LAYERED HK_NI_OFF
#endif
#if HK_NOINLINE_SOBOL
#define HK_NI_SOBOL HK_NI_ON
#else
#define HK_NI_SOBOL HK_NI_OFF
#endif
#if HK_NOINLINE_DIGITS
#define HK_NI_DIGITS HK_NI_ON
#else
#define HK_NI_DIGITS HK_NI_OFF
#endif
#if HK_NOINLINE_MISC
#define HK_NI HK_NI_ON
#else
#define HK_NI HK_NI_OFF
#endif
#define HK_PI 3.14159265358979323846f
#define HK_INF __int_as_float(0x7f800000)
#define HK_ONE_MINUS_EPS 0.99999994f

// ---- float3 ------------------------------------------------------------------------------------
HK_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
HK_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
HK_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
HK_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
HK_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
HK_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
HK_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
HK_DEV float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HK_DEV float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
HK_DEV float len3(float3 a) { return sqrtf(dot3(a, a)); }
#ifndef HK_NOINLINE_VEC
#define HK_NOINLINE_VEC 0          // norm3 (an IEEE square root and an IEEE division: ~25 instructions and two slow-path calls per call site)
#endif
#if HK_NOINLINE_VEC
static __device__ __noinline__ float3 norm3(float3 a) { float inv = 1.0f / len3(a); return f3(inv * a.x, inv * a.y, inv * a.z); }
#else
HK_DEV float3 norm3(float3 a) { float inv = 1.0f / len3(a); return f3(inv * a.x, inv * a.y, inv * a.z); }
#endif
HK_DEV float comp3(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
HK_DEV float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
HK_DEV int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
HK_DEV float lerpf(float a, float b, float t) { return (1.0f - t) * a + t * b; }
HK_DEV int floor_i(float x) { return (int)floorf(x); }
HK_DEV int trunc_i(float x) { return (int)x; }
HK_DEV int round_i(float x) { return (int)rintf(x); }

// 4x4 row-major transforms (Raycore.Transformation semantics)
HK_DEV float3 xf_point(const float* m, float3 p) {
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return f3(x, y, z);
    return f3(x / w, y / w, z / w);
}
HK_DEV float3 xf_vec(const float* m, float3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// ---- 4-wide spectrum (one float4 register quad; loads/stores are single 16-byte transactions) ---------
typedef float4 Spec;
HK_DEV Spec sp(float v) { return make_float4(v, v, v, v); }
HK_DEV Spec sp4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
HK_DEV Spec operator+(Spec a, Spec b) { return sp4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
HK_DEV Spec operator-(Spec a, Spec b) { return sp4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
HK_DEV Spec operator*(Spec a, Spec b) { return sp4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
HK_DEV Spec operator/(Spec a, Spec b) { return sp4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
HK_DEV Spec operator*(Spec a, float s) { return sp4(a.x * s, a.y * s, a.z * s, a.w * s); }
HK_DEV Spec operator*(float s, Spec a) { return a * s; }
#ifndef HK_NOINLINE_SPDIV
#define HK_NOINLINE_SPDIV 0        // Spec / float (four IEEE divisions, ~40 instructions per call site)
#endif
#if HK_NOINLINE_SPDIV
static __device__ __noinline__ Spec sp_div_f(Spec a, float s) { return sp4(a.x / s, a.y / s, a.z / s, a.w / s); }
HK_DEV Spec operator/(Spec a, float s) { return sp_div_f(a, s); }
#else
HK_DEV Spec operator/(Spec a, float s) { return sp4(a.x / s, a.y / s, a.z / s, a.w / s); }
#endif
// (a grey medium has equal coefficients at the four wavelengths: one exponential then serves all of them -- same function, same
// argument, same bits)
HK_NI Spec sp_exp(Spec a) {
    if (a.x == a.y && a.y == a.z && a.z == a.w) { const float e = dm_expf(a.x); return sp4(e, e, e, e); }
    return sp4(dm_expf(a.x), dm_expf(a.y), dm_expf(a.z), dm_expf(a.w));
}
HK_DEV Spec sp_neg(Spec a) { return sp4(-a.x, -a.y, -a.z, -a.w); }
HK_DEV Spec sp_max0(Spec a) { return sp4(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f), fmaxf(a.z, 0.0f), fmaxf(a.w, 0.0f)); }
HK_DEV float sp_avg(Spec s) { return (((s.x + s.y) + s.z) + s.w) / 4.0f; }
HK_DEV float sp_maxc(Spec s) { return fmaxf(fmaxf(fmaxf(s.x, s.y), s.z), s.w); }
HK_DEV bool sp_black(Spec s) { return s.x == 0.0f && s.y == 0.0f && s.z == 0.0f && s.w == 0.0f; }
HK_DEV float sp_get(Spec s, int i) { return i == 0 ? s.x : (i == 1 ? s.y : (i == 2 ? s.z : s.w)); }

// ---- hero wavelengths, spectral.jl:192-249 -------------------------------------------------------------
HK_DEV float visible_wavelengths_pdf(float l) {
    if (l < 360.0f || l > 830.0f) return 0.0f;
    float x = 0.0072f * (l - 538.0f);
    float c = dm_coshf(x);
    return 0.0039398042f / (c * c);
}
HK_DEV float sample_visible_wavelength(float u) { return 538.0f - 138.888889f * dm_atanhf(0.85691062f - 1.82750197f * u); }
HK_DEV void sample_wavelengths_visible(float u, float4& lambda, float4& pdf) {
    float u2 = u + 0.25f; u2 = u2 >= 1.0f ? u2 - 1.0f : u2;
    float u3 = u + 0.5f;  u3 = u3 >= 1.0f ? u3 - 1.0f : u3;
    float u4 = u + 0.75f; u4 = u4 >= 1.0f ? u4 - 1.0f : u4;
    lambda = make_float4(sample_visible_wavelength(u), sample_visible_wavelength(u2), sample_visible_wavelength(u3), sample_visible_wavelength(u4));
    pdf = make_float4(visible_wavelengths_pdf(lambda.x), visible_wavelengths_pdf(lambda.y), visible_wavelengths_pdf(lambda.z), visible_wavelengths_pdf(lambda.w));
}

// ---- hashes / RNGs ---------------------------------------------------------------------------------------
#define HK_MURMUR_M 0xc6a4a7935bd1e995ull
// MurmurHash64A specialised for the word-aligned payloads the path uses: `nwords` 32-bit words (4*nwords bytes)
template <int NW>
HK_DEV uint64_t murmur64a_words(const uint32_t* w) {
    const int r = 47;
    uint64_t h = 0ull ^ ((uint64_t)(4 * NW) * HK_MURMUR_M);
#pragma unroll
    for (int i = 0; i < NW / 2; i++) {
        uint64_t k = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        k *= HK_MURMUR_M; k ^= k >> r; k *= HK_MURMUR_M;
        h ^= k; h *= HK_MURMUR_M;
    }
    if (NW & 1) { h ^= (uint64_t)w[NW - 1]; h *= HK_MURMUR_M; }   // 4 trailing bytes
    h ^= h >> r; h *= HK_MURMUR_M; h ^= h >> r;
    return h;
}
HK_DEV uint64_t mix_bits(uint64_t v) {
    v ^= v >> 31; v *= 0x7fb5d329728ea185ull;
    v ^= v >> 27; v *= 0x81dadef4bc2dd44dull;
    v ^= v >> 33;
    return v;
}
HK_DEV uint64_t hash_f3(float3 v) { uint32_t w[3] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z)}; return murmur64a_words<3>(w); }
HK_DEV uint64_t hash_u64_f3(uint64_t s, float3 v) {
    uint32_t w[5] = {(uint32_t)s, (uint32_t)(s >> 32), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z)};
    return murmur64a_words<5>(w);
}
HK_DEV uint64_t hash_f_f2(float a, float bx, float by) { uint32_t w[3] = {__float_as_uint(a), __float_as_uint(bx), __float_as_uint(by)}; return murmur64a_words<3>(w); }
HK_DEV uint64_t hash_dim_seed(int32_t dim, uint32_t seed) { uint32_t w[2] = {(uint32_t)dim, seed}; return murmur64a_words<2>(w); }

struct Pcg32 { uint64_t state, inc; };
#define HK_PCG_MULT 0x5851f42d4c957f2dull
HK_DEV Pcg32 pcg32_init(uint64_t seq, uint64_t seed) {
    Pcg32 r; r.inc = (seq << 1) | 1ull;
    uint64_t s = r.inc;            // 0 * MULT + inc
    s += seed;
    s = s * HK_PCG_MULT + r.inc;
    r.state = s;
    return r;
}
HK_DEV uint32_t pcg32_u32(Pcg32& r) {
    uint64_t old = r.state;
    r.state = old * HK_PCG_MULT + r.inc;
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((32 - rot) & 31));
}
HK_DEV float pcg32_f32(Pcg32& r) { return fminf(HK_ONE_MINUS_EPS, (float)pcg32_u32(r) * 2.3283064e-10f); }

HK_DEV uint64_t lcg_init(float3 o, float3 d, float t_max) {   // delta-tracking.jl:28-45
    uint64_t s1 = mix_bits((uint64_t)__float_as_uint(o.x) ^ ((uint64_t)__float_as_uint(o.y) << 16) ^ ((uint64_t)__float_as_uint(o.z) << 32) ^ (uint64_t)__float_as_uint(t_max));
    uint64_t s2 = mix_bits((uint64_t)__float_as_uint(d.x) ^ ((uint64_t)__float_as_uint(d.y) << 16) ^ ((uint64_t)__float_as_uint(d.z) << 32));
    return s1 ^ s2;
}
HK_DEV float lcg_next(uint64_t& s) {                           // delta-tracking.jl:53-58
    s = s * 0x5DEECE66Dull + 11ull;
    return fminf((float)(uint32_t)(s >> 32) * 2.3283064365386963e-10f, HK_ONE_MINUS_EPS);
}

// ---- ZSobol, sobol.jl:17-309.  Integer-exact; loops are trimmed to the iterations that can contribute
// (digits i in [last_digit, n_base4_digits) and index bits below 2*n_base4_digits), which leaves every
// output bit identical to the reference's fixed 32 / 52 iteration loops. ----------------------------------
// `fast` != 0: the uploaded matrices for dimensions 0 and 1 were verified on the host (hk_upload_tables) to be the
// standard ones -- dimension 0 = bit reversal, dimension 1 = bit-reversed Pascal triangle mod 2 with columns 32.. repeating
// columns 0.. -- so sobol_bits() can use their closed forms (a GF(2)-linear map applied as a 5-level butterfly) instead of
// one dependent table load per set index bit.  Same output bits either way.
//
// Prefix cache (`top`): the base-4 digits of the Morton-indexed sample number that lie above the sample bits are a function
// of (pixel, dimension) only -- their permutations hash nothing but pixel bits -- so they are computed once per
// (resolution, seed) by k_sobol_prefix (hk_wavefront.cuh) into top[cache_slot][pixel] (u32, 4 B x pixels x dimensions in
// use; a few GB at 4K out of 180 GB HBM).  A sample then only walks the ceil(log2_spp/2) digits that depend on sample_idx
// (6 of 17 at 1080p).  Cache slots: 0,1,2 = camera dimensions 1,3,6; 3+5*depth+{0..4} = 6+7*depth+{1,3,4,6,7}.
// dimhash[cache_slot] = {u32(hash(dim+1, seed)), lo/hi of hash(dim+2, seed), 0}: the Owen-scramble seeds of zsobol_1d / _2d.
// sample_idx >= 2^log2_spp aliases into the pixel bits (reference quirk, sobol.jl:274) and takes the uncached path.
struct SobolParams {
    const uint32_t* __restrict__ M; int32_t log2_spp, n_base4_digits; uint32_t seed; int32_t fast;
    const uint32_t* __restrict__ top; const uint4* __restrict__ dimhash; uint32_t top_stride; int32_t n_top;
};
#define HK_SOBOL_SLOT_CAMERA(j) (j)                      /* j = 0,1,2 for dimensions 1,3,6 */
#define HK_SOBOL_SLOT_BOUNCE(depth, j) (3 + 5 * (depth) + (j))   /* j = 0..4 for 6+7*depth + {1,3,4,6,7} */

HK_DEV uint64_t left_shift2(uint64_t x) {
    x &= 0xffffffffull;
    x = (x ^ (x << 16)) & 0x0000ffff0000ffffull;
    x = (x ^ (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x ^ (x << 2)) & 0x3333333333333333ull;
    x = (x ^ (x << 1)) & 0x5555555555555555ull;
    return x;
}
HK_DEV uint64_t encode_morton2(uint32_t x, uint32_t y) { return (left_shift2(y) << 1) | left_shift2(x); }
HK_DEV uint32_t fast_owen_scramble(uint32_t v, uint32_t seed) {
    v = __brev(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return __brev(v);
}
// the 24 permutations of (0,1,2,3), 2 bits per entry packed into one byte each (sobol.jl:155-180), held as six words in
// registers and indexed with PRMT: a per-lane index into __constant__ memory serialises over the distinct addresses of a
// warp (it was the top stall line of k_shade in ncu), three byte-permutes and two selects do not.
HK_DEV uint32_t perm4_byte(uint32_t p) {   // p in [0, 24)
    const uint32_t a = __byte_perm(0x78D8B4E4u, 0xB1E19C6Cu, p & 7u);      // entries 0..7   (selector uses the low 3 bits)
    const uint32_t b = __byte_perm(0x8D2D39C9u, 0x72D236C6u, p & 7u);      // entries 8..15
    const uint32_t c = __byte_perm(0x87271E4Eu, 0x93634B1Bu, p & 7u);      // entries 16..23
    return (p < 8u ? a : (p < 16u ? b : c)) & 0xFFu;
}
// not inlined: the digit loop is ~600 instructions and has five call sites per shading kernel; inlining all of them made
// the kernels instruction-fetch bound (stall_no_instruction was the top stall reason)
// digits i_hi .. i_lo (inclusive, i_lo >= log2_spp & 1) of zsobol_get_sample_index, sobol.jl:269-291
HK_NI_DIGITS uint64_t zsobol_digits(uint64_t morton, uint64_t dmix, int pow2, int i_hi, int i_lo) {
    uint64_t idx = 0;
    for (int i = i_hi; i >= i_lo; --i) {
        int shift = 2 * i - pow2;
        uint32_t digit = (uint32_t)(morton >> shift) & 3u;
        int hs = shift + 2;
        uint64_t higher = hs >= 64 ? 0ull : (morton >> hs);
        // (h >> 24) % 24 for the 40-bit value h >> 24, in 32-bit arithmetic: 2^24 mod 24 = 16, so
        // h>>24 = top16 * 2^24 + low24  ==  top16 * 16 + low24  (mod 24), which is < 2^25
        const uint64_t hm = mix_bits(higher ^ dmix);
        const uint32_t p = ((uint32_t)(hm >> 48) * 16u + ((uint32_t)(hm >> 24) & 0xFFFFFFu)) % 24u;
        const uint64_t pd = (perm4_byte(p) >> (2 * digit)) & 3u;
        idx |= pd << shift;
    }
    return idx;
}
HK_DEV uint64_t zsobol_dmix(int32_t dim) { return 0x55555555ull * (uint64_t)(int64_t)dim; }
HK_DEV uint64_t zsobol_sample_index(uint64_t morton, int32_t dim, int32_t log2_spp, int32_t nb4) {
    const int pow2 = log2_spp & 1;
    const uint64_t dmix = zsobol_dmix(dim);
    uint64_t idx = zsobol_digits(morton, dmix, pow2, nb4 - 1, pow2);
    if (pow2) {
        uint64_t digit = morton & 1ull;
        idx |= digit ^ (mix_bits((morton >> 1) ^ dmix) & 1ull);
    }
    return idx;
}
// first digit whose permutation depends on pixel bits only
HK_DEV int zsobol_first_pixel_digit(int32_t log2_spp) { return (log2_spp + 1) >> 1; }
// value stored in the prefix cache for one (pixel, dimension)
HK_DEV uint32_t zsobol_prefix(uint32_t px, uint32_t py, int32_t dim, int32_t log2_spp, int32_t nb4) {
    const int pow2 = log2_spp & 1, ilo = zsobol_first_pixel_digit(log2_spp);
    const uint64_t morton = encode_morton2(px, py) << log2_spp;
    return (uint32_t)(zsobol_digits(morton, zsobol_dmix(dim), pow2, nb4 - 1, ilo) >> (2 * ilo - pow2));
}
HK_DEV uint32_t sobol_bits(uint64_t a, int32_t dimension, const uint32_t* __restrict__ M) {
    uint32_t v = 0;
    const uint32_t* m = M + dimension * 52;
    for (int bit = 0; a != 0; ++bit, a >>= 1) if (a & 1) v ^= __ldg(m + bit);
    return v;
}
HK_DEV uint32_t pascal_butterfly(uint32_t x) {   // y_i = XOR_{j >= i} C(j, i) x_j  (Kronecker power of [[1,1],[0,1]])
    x ^= (x >> 1) & 0x55555555u; x ^= (x >> 2) & 0x33333333u; x ^= (x >> 4) & 0x0f0f0f0fu;
    x ^= (x >> 8) & 0x00ff00ffu; x ^= (x >> 16) & 0x0000ffffu;
    return x;
}
HK_DEV uint32_t sobol_bits0(const SobolParams& S, uint64_t a) { return S.fast ? __brev((uint32_t)a) : sobol_bits(a, 0, S.M); }
HK_DEV uint32_t sobol_bits1(const SobolParams& S, uint64_t a) {
    return S.fast ? __brev(pascal_butterfly((uint32_t)a) ^ pascal_butterfly((uint32_t)(a >> 32))) : sobol_bits(a, 1, S.M);
}
HK_DEV float sobol_to_float(uint32_t v) { return fminf((float)v * 2.3283064365386963e-10f, HK_ONE_MINUS_EPS); }
// Morton-indexed sample number of (pixel, sample_idx) in one dimension + the Owen-scramble seeds of that dimension.
// cslot < 0 (or no cache uploaded): everything is computed from scratch -- same bits either way.
struct ZSample { uint64_t si; uint32_t h1, h2lo, h2hi; };
HK_DEV ZSample zsobol_index(const SobolParams& S, int32_t px, int32_t py, int32_t sample_idx, int32_t dim, int32_t cslot, uint32_t pix) {
    ZSample z;
    const uint64_t morton = (encode_morton2((uint32_t)px, (uint32_t)py) << S.log2_spp) | (uint64_t)(int64_t)sample_idx;
    if (S.top != nullptr && cslot >= 0 && cslot < S.n_top && ((uint32_t)sample_idx >> S.log2_spp) == 0u) {
        const int pow2 = S.log2_spp & 1, ilo = zsobol_first_pixel_digit(S.log2_spp);
        const uint64_t dmix = zsobol_dmix(dim);
        uint64_t idx = ((uint64_t)__ldg(S.top + (size_t)cslot * S.top_stride + pix) << (2 * ilo - pow2)) | zsobol_digits(morton, dmix, pow2, ilo - 1, pow2);
        if (pow2) idx |= (morton & 1ull) ^ (mix_bits((morton >> 1) ^ dmix) & 1ull);
        z.si = idx;
        const uint4 h = __ldg(S.dimhash + cslot);
        z.h1 = h.x; z.h2lo = h.y; z.h2hi = h.z;
    } else {
        z.si = zsobol_sample_index(morton, dim, S.log2_spp, S.n_base4_digits);
        z.h1 = (uint32_t)hash_dim_seed(dim + 1, S.seed);
        const uint64_t b2 = hash_dim_seed(dim + 2, S.seed);
        z.h2lo = (uint32_t)b2; z.h2hi = (uint32_t)(b2 >> 32);
    }
    return z;
}
HK_NI_SOBOL float zsobol_1d(const SobolParams& S, int32_t px, int32_t py, int32_t sample_idx, int32_t dim, int32_t cslot = -1, uint32_t pix = 0) {
    const ZSample z = zsobol_index(S, px, py, sample_idx, dim, cslot, pix);
    return sobol_to_float(fast_owen_scramble(sobol_bits0(S, z.si), z.h1));
}
HK_NI_SOBOL float2 zsobol_2d(const SobolParams& S, int32_t px, int32_t py, int32_t sample_idx, int32_t dim, int32_t cslot = -1, uint32_t pix = 0) {
    const ZSample z = zsobol_index(S, px, py, sample_idx, dim, cslot, pix);
    return make_float2(sobol_to_float(fast_owen_scramble(sobol_bits0(S, z.si), z.h2lo)),
                       sobol_to_float(fast_owen_scramble(sobol_bits1(S, z.si), z.h2hi)));
}

// ---- sampling primitives, sampling.jl:5-33, spectral-eval.jl:3514-3533 ------------------------------------
HK_DEV float2 concentric_sample_disk(float2 u) {
    float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
    bool xl = fabsf(ox) > fabsf(oy);
    float r = xl ? ox : oy;
    float th = xl ? (oy / (ox + 1.0e-10f)) * HK_PI / 4.0f : HK_PI / 2.0f - (ox / (oy + 1.0e-10f)) * HK_PI / 4.0f;
    return make_float2(r * dm_cosf(th), r * dm_sinf(th));
}
HK_DEV float3 cosine_sample_hemisphere(float2 u) {
    float2 d = concentric_sample_disk(u);
    return f3(d.x, d.y, sqrtf(fmaxf(0.0f, 1.0f - d.x * d.x - d.y * d.y)));
}
struct Frame { float3 t, b, n; };
HK_DEV Frame make_frame(float3 n) {
    Frame f; f.n = n;
    if (fabsf(n.x) > fabsf(n.y)) { float inv = 1.0f / sqrtf(n.x * n.x + n.z * n.z); f.t = f3(n.z * inv, 0.0f, -n.x * inv); }
    else { float inv = 1.0f / sqrtf(n.y * n.y + n.z * n.z); f.t = f3(0.0f, n.z * inv, -n.y * inv); }
    f.b = cross3(n, f.t);
    return f;
}
HK_DEV float3 to_world(const Frame& f, float3 l) { return f.t * l.x + f.b * l.y + f.n * l.z; }
HK_DEV float3 to_local(const Frame& f, float3 v) { return f3(dot3(v, f.t), dot3(v, f.b), dot3(v, f.n)); }
HK_DEV float3 reflect3(float3 wo, float3 n) { return -wo + 2.0f * dot3(wo, n) * n; }
