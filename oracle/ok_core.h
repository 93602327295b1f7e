// ok_core.h — ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference's vector / spectrum / hash / RNG / sampler primitives.
// Compile with -ffp-contract=off: Julia does not contract a*b+c to FMA without @fastmath/muladd.
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>
#include "../hikari_jl_b200/csrc/hk_detmath.h"   // the ONE f32 libm shared with the CUDA library (bit-identical transcendentals)

namespace ok {

static constexpr float PI_F = 3.14159265358979323846f;  // Float32(π)
static constexpr float INF_F = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------------------------
// Vec3 (GeometryBasics Vec3f / Point3f; StaticArrays semantics: dot = ((x*x)+(y*y))+(z*z),
// normalize(v) = inv(norm(v)) * v)
// ---------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(V3 a, V3 b) { return !(a == b); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { float inv = 1.0f / norm(a); return V3(inv * a.x, inv * a.y, inv * a.z); }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// src/spectrum.jl:33  lerp(v1,v2,t) = (1 - t) * v1 + t * v2
inline float lerpf(float v1, float v2, float t) { return (1.0f - t) * v1 + t * v2; }

struct V2 { float x, y; V2() : x(0), y(0) {} V2(float a, float b) : x(a), y(b) {} };

// src/materials/bsdf.jl:126-135  unsafe_trunc helpers
inline int32_t u_int32(float x) { return (int32_t)x; }
inline int32_t floor_int32(float x) { return (int32_t)std::floor(x); }
inline int32_t round_int32(float x) { return (int32_t)std::nearbyintf(x); }  // Julia round = ties-to-even

// 4x4 row-major transform application (Raycore.Transformation semantics)
inline V3 xform_point(const float* m, V3 p) {
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return V3(x, y, z);
    return V3(x / w, y / w, z / w);
}
inline V3 xform_vec(const float* m, V3 v) {
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z,
              m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// ---------------------------------------------------------------------------------------------
// SampledSpectrum{4} / SampledWavelengths{4}   src/spectral/spectral.jl:10-126
// ---------------------------------------------------------------------------------------------
struct Spec {
    float v[4];
    Spec() { v[0] = v[1] = v[2] = v[3] = 0.0f; }
    explicit Spec(float a) { v[0] = v[1] = v[2] = v[3] = a; }
    Spec(float a, float b, float c, float d) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    float operator[](int i) const { return v[i]; }
};
#define OK_SPEC_OP(op) \
    inline Spec operator op(const Spec& a, const Spec& b) { return Spec(a.v[0] op b.v[0], a.v[1] op b.v[1], a.v[2] op b.v[2], a.v[3] op b.v[3]); }
OK_SPEC_OP(+) OK_SPEC_OP(-) OK_SPEC_OP(*) OK_SPEC_OP(/)
#undef OK_SPEC_OP
inline Spec operator*(const Spec& a, float s) { return Spec(a.v[0] * s, a.v[1] * s, a.v[2] * s, a.v[3] * s); }
inline Spec operator*(float s, const Spec& a) { return a * s; }   // spectral.jl:46  s*a = a*s
inline Spec operator/(const Spec& a, float s) { return Spec(a.v[0] / s, a.v[1] / s, a.v[2] / s, a.v[3] / s); }
inline Spec operator-(const Spec& a) { return Spec(-a.v[0], -a.v[1], -a.v[2], -a.v[3]); }
inline Spec exp(const Spec& a) { return Spec(dm_expf(a.v[0]), dm_expf(a.v[1]), dm_expf(a.v[2]), dm_expf(a.v[3])); }
// spectral.jl:63-65  sum(s.data)/N ; Julia sum of a 4-tuple = ((a+b)+c)+d
inline float average(const Spec& s) { return (((s.v[0] + s.v[1]) + s.v[2]) + s.v[3]) / 4.0f; }
inline float max_component(const Spec& s) { return std::max(std::max(std::max(s.v[0], s.v[1]), s.v[2]), s.v[3]); }
inline bool is_black(const Spec& s) { return s.v[0] == 0.0f && s.v[1] == 0.0f && s.v[2] == 0.0f && s.v[3] == 0.0f; }

struct Wavelengths { float lambda[4]; float pdf[4]; };

// spectral.jl:192-200
inline float visible_wavelengths_pdf(float lambda) {
    if (lambda < 360.0f || lambda > 830.0f) return 0.0f;
    float x = 0.0072f * (lambda - 538.0f);
    float c = dm_coshf(x);
    return 0.0039398042f / (c * c);
}
// spectral.jl:210-213
inline float sample_visible_wavelengths(float u) {
    return 538.0f - 138.888889f * dm_atanhf(0.85691062f - 1.82750197f * u);
}
// spectral.jl:221-249
inline Wavelengths sample_wavelengths_visible(float u) {
    Wavelengths w;
    float us[4];
    us[0] = u;
    float u2 = u + 0.25f; us[1] = u2 >= 1.0f ? u2 - 1.0f : u2;
    float u3 = u + 0.5f;  us[2] = u3 >= 1.0f ? u3 - 1.0f : u3;
    float u4 = u + 0.75f; us[3] = u4 >= 1.0f ? u4 - 1.0f : u4;
    for (int i = 0; i < 4; i++) w.lambda[i] = sample_visible_wavelengths(us[i]);
    for (int i = 0; i < 4; i++) w.pdf[i] = visible_wavelengths_pdf(w.lambda[i]);
    return w;
}

// ---------------------------------------------------------------------------------------------
// Hashes and RNGs   src/materials/spectral-eval.jl:575-815, src/integrators/volpath/delta-tracking.jl:28-58
// ---------------------------------------------------------------------------------------------
// spectral-eval.jl:575-633  MurmurHash64A over a byte buffer
inline uint64_t murmur_hash_64a(const uint8_t* data, size_t n, uint64_t seed = 0) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ ((uint64_t)n * m);
    size_t n_chunks = n / 8;
    for (size_t i = 0; i < n_chunks; i++) {
        uint64_t k = 0;
        for (int b = 0; b < 8; b++) k |= (uint64_t)data[8 * i + b] << (8 * b);
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    size_t rem = n & 7, off = 8 * n_chunks;
    if (rem >= 7) h ^= (uint64_t)data[off + 6] << 48;
    if (rem >= 6) h ^= (uint64_t)data[off + 5] << 40;
    if (rem >= 5) h ^= (uint64_t)data[off + 4] << 32;
    if (rem >= 4) h ^= (uint64_t)data[off + 3] << 24;
    if (rem >= 3) h ^= (uint64_t)data[off + 2] << 16;
    if (rem >= 2) h ^= (uint64_t)data[off + 1] << 8;
    if (rem >= 1) { h ^= (uint64_t)data[off]; h *= m; }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
// spectral-eval.jl:641-648
inline uint64_t mix_bits(uint64_t v) {
    v ^= v >> 31; v *= 0x7fb5d329728ea185ull;
    v ^= v >> 27; v *= 0x81dadef4bc2dd44dull;
    v ^= v >> 33;
    return v;
}
inline void put_f32(uint8_t* b, float f) { uint32_t u; std::memcpy(&u, &f, 4); for (int i = 0; i < 4; i++) b[i] = (uint8_t)(u >> (8 * i)); }
inline void put_u64(uint8_t* b, uint64_t u) { for (int i = 0; i < 8; i++) b[i] = (uint8_t)(u >> (8 * i)); }
// spectral-eval.jl:690-741  pbrt_hash overloads
inline uint64_t pbrt_hash(float v) { uint8_t b[4]; put_f32(b, v); return murmur_hash_64a(b, 4); }
inline uint64_t pbrt_hash(V3 v) { uint8_t b[12]; put_f32(b, v.x); put_f32(b + 4, v.y); put_f32(b + 8, v.z); return murmur_hash_64a(b, 12); }
inline uint64_t pbrt_hash(uint64_t s, V3 v) { uint8_t b[20]; put_u64(b, s); put_f32(b + 8, v.x); put_f32(b + 12, v.y); put_f32(b + 16, v.z); return murmur_hash_64a(b, 20); }
inline uint64_t pbrt_hash(uint64_t a, float f) { uint8_t b[12]; put_u64(b, a); put_f32(b + 8, f); return murmur_hash_64a(b, 12); }
inline uint64_t pbrt_hash(float a, V2 p) { uint8_t b[12]; put_f32(b, a); put_f32(b + 4, p.x); put_f32(b + 8, p.y); return murmur_hash_64a(b, 12); }

// spectral-eval.jl:745-815  PCG32
struct PCG32 { uint64_t state, inc; };
static constexpr uint64_t PCG32_MULT = 0x5851f42d4c957f2dull;
inline PCG32 pcg32_init(uint64_t seq_index, uint64_t seed) {
    PCG32 r; r.inc = (seq_index << 1) | 1ull;
    uint64_t s = 0;
    s = s * PCG32_MULT + r.inc;
    s += seed;
    s = s * PCG32_MULT + r.inc;
    r.state = s;
    return r;
}
inline uint32_t pcg32_u32(PCG32& r) {
    uint64_t old = r.state;
    r.state = old * PCG32_MULT + r.inc;
    uint32_t xs = (uint32_t)((((old >> 18) ^ old) >> 27) & 0xFFFFFFFFull);
    uint32_t rot = (uint32_t)((old >> 59) & 0x1F);
    return (xs >> rot) | (xs << ((32 - rot) & 31));
}
static constexpr float ONE_MINUS_EPS = 0.99999994f;   // Float32(1) - eps(Float32)
inline float pcg32_f32(PCG32& r) {
    uint32_t u = pcg32_u32(r);
    return std::min(ONE_MINUS_EPS, (float)u * 2.3283064e-10f);
}

// ---------------------------------------------------------------------------------------------
// ZSobol sampler   src/sampler/sobol.jl
// ---------------------------------------------------------------------------------------------
// sobol.jl:17-31
inline uint64_t zsobol_hash(int32_t dimension, uint32_t seed) {
    uint8_t b[8]; uint32_t d = (uint32_t)dimension;
    for (int i = 0; i < 4; i++) { b[i] = (uint8_t)(d >> (8 * i)); b[4 + i] = (uint8_t)(seed >> (8 * i)); }
    return murmur_hash_64a(b, 8, 0);
}
// sobol.jl:42-60
inline uint64_t left_shift2(uint64_t x) {
    x &= 0xffffffffull;
    x = (x ^ (x << 16)) & 0x0000ffff0000ffffull;
    x = (x ^ (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x ^ (x << 2)) & 0x3333333333333333ull;
    x = (x ^ (x << 1)) & 0x5555555555555555ull;
    return x;
}
inline uint64_t encode_morton2(uint32_t x, uint32_t y) { return (left_shift2(y) << 1) | left_shift2(x); }
inline uint32_t bitreverse32(uint32_t n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    return n;
}
// sobol.jl:72-80
inline uint32_t fast_owen_scramble(uint32_t v, uint32_t seed) {
    v = bitreverse32(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return bitreverse32(v);
}
// sobol.jl:108-127
inline float sobol_sample(int64_t a, int32_t dimension, uint32_t scramble_seed, const uint32_t* M) {
    uint32_t v = 0;
    int base = dimension * 52;
    for (int bit = 0; bit < 52; bit++) {
        uint32_t mask = (uint32_t)((a >> bit) & 1) * 0xffffffffu;
        v ^= M[base + bit] & mask;
    }
    v = fast_owen_scramble(v, scramble_seed);
    return std::min((float)v * 2.3283064365386963e-10f, ONE_MINUS_EPS);
}
// sobol.jl:155-180
static const uint8_t PERMUTATIONS_4WAY[24][4] = {
    {0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 1, 3}, {0, 2, 3, 1}, {0, 3, 2, 1}, {0, 3, 1, 2},
    {1, 0, 2, 3}, {1, 0, 3, 2}, {1, 2, 0, 3}, {1, 2, 3, 0}, {1, 3, 2, 0}, {1, 3, 0, 2},
    {2, 1, 0, 3}, {2, 1, 3, 0}, {2, 0, 1, 3}, {2, 0, 3, 1}, {2, 3, 0, 1}, {2, 3, 1, 0},
    {3, 1, 2, 0}, {3, 1, 0, 2}, {3, 2, 1, 0}, {3, 2, 0, 1}, {3, 0, 2, 1}, {3, 0, 1, 2}};
// sobol.jl:211-258 (literal restatement, including the 32 fixed iterations)
inline uint64_t zsobol_get_sample_index(uint64_t morton, int32_t dimension, int32_t log2_spp, int32_t n_base4_digits) {
    uint64_t sample_index = 0;
    int32_t pow2_flag = log2_spp & 1;
    int32_t last_digit = pow2_flag, pow2_adjust = pow2_flag;
    for (int32_t iter0 = 0; iter0 < 32; iter0++) {
        int32_t i = n_base4_digits - 1 - iter0;
        int32_t raw_shift = 2 * i - pow2_adjust;
        int32_t digit_shift = raw_shift > 0 ? raw_shift : 0;
        int32_t digit = (int32_t)((morton >> digit_shift) & 3ull);
        int sh = digit_shift + 2;
        uint64_t higher = sh >= 64 ? 0ull : (morton >> sh);
        uint64_t hv = mix_bits(higher ^ (0x55555555ull * (uint64_t)(int64_t)dimension));
        int32_t p = (int32_t)((hv >> 24) % 24ull);
        uint64_t permuted = PERMUTATIONS_4WAY[p][digit];
        if (i >= last_digit) sample_index |= permuted << digit_shift;
    }
    if (pow2_flag) {
        uint64_t digit = morton & 1ull;
        uint64_t xor_bit = mix_bits((morton >> 1) ^ (0x55555555ull * (uint64_t)(int64_t)dimension)) & 1ull;
        sample_index |= (digit ^ xor_bit);
    }
    return sample_index;
}
struct SobolRNG { const uint32_t* M; int32_t log2_spp, n_base4_digits; uint32_t seed; int32_t width; };
// sobol.jl:269-282
inline float zsobol_1d(const SobolRNG& r, int32_t px, int32_t py, int32_t sample_idx, int32_t dim) {
    uint64_t morton = (encode_morton2((uint32_t)px, (uint32_t)py) << r.log2_spp) | (uint64_t)(int64_t)sample_idx;
    uint64_t si = zsobol_get_sample_index(morton, dim, r.log2_spp, r.n_base4_digits);
    uint32_t h = (uint32_t)zsobol_hash(dim + 1, r.seed);
    return sobol_sample((int64_t)si, 0, h, r.M);
}
// sobol.jl:290-309
inline V2 zsobol_2d(const SobolRNG& r, int32_t px, int32_t py, int32_t sample_idx, int32_t dim) {
    uint64_t morton = (encode_morton2((uint32_t)px, (uint32_t)py) << r.log2_spp) | (uint64_t)(int64_t)sample_idx;
    uint64_t si = zsobol_get_sample_index(morton, dim, r.log2_spp, r.n_base4_digits);
    uint64_t bits = zsobol_hash(dim + 2, r.seed);
    return V2(sobol_sample((int64_t)si, 0, (uint32_t)bits, r.M), sobol_sample((int64_t)si, 1, (uint32_t)(bits >> 32), r.M));
}

// ---------------------------------------------------------------------------------------------
// Sampling primitives   src/sampler/sampling.jl:5-33
// ---------------------------------------------------------------------------------------------
inline V2 concentric_sample_disk(V2 u) {
    float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
    float ax = std::fabs(ox), ay = std::fabs(oy);
    float sx = ox + 1.0e-10f, sy = oy + 1.0e-10f;
    bool xl = ax > ay;
    float r = xl ? ox : oy;
    float th = xl ? (oy / sx) * PI_F / 4.0f : PI_F / 2.0f - (ox / sy) * PI_F / 4.0f;
    return V2(r * dm_cosf(th), r * dm_sinf(th));
}
inline V3 cosine_sample_hemisphere(V2 u) {
    V2 d = concentric_sample_disk(u);
    float z = std::sqrt(std::max(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return V3(d.x, d.y, z);
}

// src/materials/spectral-eval.jl:3514-3533, 3579-3581
inline void coordinate_system(V3 n, V3& t, V3& b) {
    if (std::fabs(n.x) > std::fabs(n.y)) {
        float inv = 1.0f / std::sqrt(n.x * n.x + n.z * n.z);
        t = V3(n.z * inv, 0.0f, -n.x * inv);
    } else {
        float inv = 1.0f / std::sqrt(n.y * n.y + n.z * n.z);
        t = V3(0.0f, n.z * inv, -n.y * inv);
    }
    b = cross(n, t);
}
inline V3 local_to_world(V3 l, V3 n, V3 t, V3 b) { return t * l.x + b * l.y + n * l.z; }
inline V3 world_to_local(V3 v, V3 n, V3 t, V3 b) { return V3(dot(v, t), dot(v, b), dot(v, n)); }
// spectral-eval.jl:1127-1129
inline V3 reflect(V3 wo, V3 n) { return -wo + 2.0f * dot(wo, n) * n; }

}  // namespace ok
