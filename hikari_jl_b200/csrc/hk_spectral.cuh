// hk_spectral.cuh — device scene tables + RGB->spectrum lookup, uplifts, D65, CIE XYZ, filter and camera.
// Reference: src/spectral/rgb2spec.jl:83-167, uplift.jl:255-308,437-538, color.jl:364-440,572-579,
// src/filter.jl:727-953, src/camera/perspective.jl:95-128, src/spectral/piecewise-linear.jl:11-31.
#pragma once
#include "hk_math.cuh"
#include "../../include/hikari_cuda.h"

// The sRGB coefficient table is re-laid-out at upload from the reference's [coef][x][y][z][maxc] memory order
// to [maxc][z][y][x] float4 (c0,c1,c2,pad): the 8 trilinear corners become 8 x 16-byte loads in 4 runs of
// 32 contiguous bytes instead of 24 scattered 4-byte gathers.  Values are bit-identical.
struct DevTables {
    const uint32_t* __restrict__ sobol;
    const float* __restrict__ cie_x; const float* __restrict__ cie_y; const float* __restrict__ cie_z;
    const float* __restrict__ d65;
    const float* __restrict__ rgb_scale;
    const float4* __restrict__ rgb_coeffs;   // [3][res][res][res]
    int32_t rgb_res;
    // Uplift cache: rgb_to_spectrum (clamp, 64-entry search, 8-corner trilinear of 3 coefficients) of every CONSTANT colour of
    // the scene -- material parameters, light spectra, medium coefficients -- is evaluated once per upload by
    // k_precompute_uplifts with these very device functions, as (c0, c1, c2, scale); shading evaluates the cached
    // polynomial at the path's wavelengths.  Same bits as the per-hit evaluation, ~10 % fewer shading instructions and no
    // dependent table loads.  Null pointers = not built yet: everything falls back to the direct evaluation.
    const float4* __restrict__ mat_pre;   const void* mat_base;     // [2 * n_materials]: slot A, slot B (per type, hk_bsdf.cuh)
    const float4* __restrict__ light_pre; const void* light_base;   // [2 * n_lights]: illuminant spectrum, area-light Le
    const float4* __restrict__ med_pre;   const void* med_base;     // [3 * n_media]: sigma_a, sigma_s, Le
};

struct Poly3 { float c0, c1, c2; };
// HK_SPEC_FN: the uplift helpers are called from a dozen places per shading kernel; as real functions (not inlined)
// the kernels shrink by a third, which matters because they were instruction-fetch bound (stall_no_instruction)
#ifndef HK_NOINLINE_UPLIFT
#define HK_NOINLINE_UPLIFT 1
#endif
#if HK_NOINLINE_UPLIFT
#define HK_SPEC_FN static __device__ __noinline__
#else
#define HK_SPEC_FN HK_DEV
#endif
HK_DEV float sigmoidf_(float x) {
    if (isinf(x)) return x > 0 ? 1.0f : 0.0f;
    return 0.5f + x / (2.0f * sqrtf(1.0f + x * x));
}
HK_DEV float poly_eval(const Poly3& p, float l) { return sigmoidf_(p.c0 * l * l + p.c1 * l + p.c2); }
HK_SPEC_FN float poly_max_value(Poly3 p) {
    float r = fmaxf(poly_eval(p, 360.0f), poly_eval(p, 830.0f));
    if (p.c0 != 0.0f) {
        float lc = -p.c1 / (2.0f * p.c0);
        if (360.0f <= lc && lc <= 830.0f) r = fmaxf(r, poly_eval(p, lc));
    }
    return r;
}
// (arguments by value: a reference parameter of a real function would push the caller's struct to local memory)
HK_SPEC_FN Poly3 rgb_to_spectrum_tab(const float* __restrict__ rgb_scale, const float4* __restrict__ rgb_coeffs, int res, float r, float g, float b) {
    r = clampf(r, 0.0f, 1.0f); g = clampf(g, 0.0f, 1.0f); b = clampf(b, 0.0f, 1.0f);
    if (r == g && g == b) {
        float c2 = (r > 0.0f && r < 1.0f) ? (r - 0.5f) / sqrtf(r * (1.0f - r)) : (r <= 0.0f ? -1.0e10f : 1.0e10f);
        return Poly3{0.0f, 0.0f, c2};
    }
    int maxc = r > g ? (r > b ? 0 : 2) : (g > b ? 1 : 2);
    float z = maxc == 0 ? r : (maxc == 1 ? g : b);
    float xc = maxc == 0 ? g : (maxc == 1 ? b : r);
    float yc = maxc == 0 ? b : (maxc == 1 ? r : g);
    float x = xc * (float)(res - 1) / z;
    float y = yc * (float)(res - 1) / z;
    // last i in [1, res-1] with scale[i] < z (scale is increasing), 1 if none — same result as the linear scan
    int zi = 1;
    { int lo = 1, hi = res - 1; while (lo <= hi) { int mid = (lo + hi) >> 1; if (__ldg(rgb_scale + mid - 1) < z) { zi = mid; lo = mid + 1; } else hi = mid - 1; } }
    zi = min(zi, res - 1);
    int xi = min(trunc_i(x) + 1, res - 1), yi = min(trunc_i(y) + 1, res - 1);
    float dx = x - (float)(xi - 1), dy = y - (float)(yi - 1);
    float s0 = __ldg(rgb_scale + zi - 1), s1 = __ldg(rgb_scale + zi);
    float dz = (z - s0) / (s1 - s0);
    const float4* base = rgb_coeffs + (((size_t)maxc * res + (zi - 1)) * res + (yi - 1)) * res + (xi - 1);
    const size_t sy = (size_t)res, sz = (size_t)res * res;
    float4 c000 = __ldg(base), c001 = __ldg(base + 1), c010 = __ldg(base + sy), c011 = __ldg(base + sy + 1);
    float4 c100 = __ldg(base + sz), c101 = __ldg(base + sz + 1), c110 = __ldg(base + sz + sy), c111 = __ldg(base + sz + sy + 1);
    float mx = 1.0f - dx, my = 1.0f - dy, mz = 1.0f - dz;
#define HK_TRI(f) (mz * (my * (mx * c000.f + dx * c001.f) + dy * (mx * c010.f + dx * c011.f)) + dz * (my * (mx * c100.f + dx * c101.f) + dy * (mx * c110.f + dx * c111.f)))
    Poly3 p{HK_TRI(x), HK_TRI(y), HK_TRI(z)};
#undef HK_TRI
    return p;
}
HK_DEV Poly3 rgb_to_spectrum(const DevTables& T, float r, float g, float b) { return rgb_to_spectrum_tab(T.rgb_scale, T.rgb_coeffs, T.rgb_res, r, g, b); }
HK_SPEC_FN Spec poly_eval4(Poly3 p, float4 l) { return sp4(poly_eval(p, l.x), poly_eval(p, l.y), poly_eval(p, l.z), poly_eval(p, l.w)); }
HK_DEV Spec uplift_rgb(const DevTables& T, float r, float g, float b, float4 lambda) { return poly_eval4(rgb_to_spectrum(T, r, g, b), lambda); }
HK_DEV Spec uplift_rgb_unbounded(const DevTables& T, float r, float g, float b, float4 lambda) {
    float m = fmaxf(fmaxf(r, g), b);
    if (m <= 0.0f) return sp(0.0f);
    Poly3 p = rgb_to_spectrum(T, r / m, g / m, b / m);
    float s = m / poly_max_value(p);
    return poly_eval4(p, lambda) * s;
}
// cached forms: make_* run at upload (and as the fallback), pre_* at every hit
HK_DEV float4 make_pre_bounded(const DevTables& T, float r, float g, float b) { Poly3 p = rgb_to_spectrum(T, r, g, b); return make_float4(p.c0, p.c1, p.c2, 1.0f); }
HK_DEV float4 make_pre_unbounded(const DevTables& T, float r, float g, float b) {
    float m = fmaxf(fmaxf(r, g), b);
    if (m <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    Poly3 p = rgb_to_spectrum(T, r / m, g / m, b / m);
    return make_float4(p.c0, p.c1, p.c2, m / poly_max_value(p));
}
HK_DEV float4 make_pre_illuminant(const DevTables& T, float r, float g, float b) {
    float m = fmaxf(fmaxf(r, g), b);
    if (m <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float s = 2.0f * m;
    Poly3 p = rgb_to_spectrum(T, r / s, g / s, b / s);
    return make_float4(p.c0, p.c1, p.c2, s);
}
HK_DEV Spec pre_bounded(float4 q, float4 lam) { return poly_eval4(Poly3{q.x, q.y, q.z}, lam); }
HK_DEV Spec pre_unbounded(float4 q, float4 lam) { if (q.w == 0.0f) return sp(0.0f); return poly_eval4(Poly3{q.x, q.y, q.z}, lam) * q.w; }
HK_DEV float sample_d65(const float* __restrict__ d65, float l) {
    if (l <= 300.0f) return __ldg(d65);
    if (l >= 830.0f) return __ldg(d65 + 106);
    float t = (l - 300.0f) / 5.0f;
    int fl = floor_i(t);
    int idx = clampi(fl + 1, 1, 106);
    float fr = t - (float)fl;
    return __ldg(d65 + idx - 1) * (1.0f - fr) + __ldg(d65 + idx) * fr;
}
HK_SPEC_FN Spec illuminant_eval_tab(const float* __restrict__ d65, Poly3 p, float scale, float4 l) {
    return sp4(scale * poly_eval(p, l.x) * sample_d65(d65, l.x), scale * poly_eval(p, l.y) * sample_d65(d65, l.y),
               scale * poly_eval(p, l.z) * sample_d65(d65, l.z), scale * poly_eval(p, l.w) * sample_d65(d65, l.w));
}
HK_DEV Spec illuminant_eval(const DevTables& T, const Poly3& p, float scale, float4 l) { return illuminant_eval_tab(T.d65, p, scale, l); }
HK_DEV Spec pre_illuminant(const DevTables& T, float4 q, float4 lam) { if (q.w == 0.0f) return sp(0.0f); return illuminant_eval(T, Poly3{q.x, q.y, q.z}, q.w, lam); }
HK_DEV Spec uplift_rgb_illuminant(const DevTables& T, float r, float g, float b, float4 lambda) {
    float m = fmaxf(fmaxf(r, g), b);
    if (m <= 0.0f) return sp(0.0f);
    float s = 2.0f * m;
    return illuminant_eval(T, rgb_to_spectrum(T, r / s, g / s, b / s), s, lambda);
}
HK_DEV float pls_sample(const float* __restrict__ lam, const float* __restrict__ val, int N, float l) {
    if (l <= __ldg(lam)) return __ldg(val);
    if (l >= __ldg(lam + N - 1)) return __ldg(val + N - 1);
    int lo = 1, hi = N;
    while (lo + 1 < hi) { int mid = (lo + hi) >> 1; if (__ldg(lam + mid - 1) <= l) lo = mid; else hi = mid; }
    float l0 = __ldg(lam + lo - 1), l1 = __ldg(lam + hi - 1);
    float t = (l - l0) / (l1 - l0);
    return __ldg(val + lo - 1) * (1.0f - t) + __ldg(val + hi - 1) * t;
}
HK_DEV float cie_lookup(const float* __restrict__ tab, float l) {
    int off = round_i(l) - 360;
    return (off < 0 || off >= 471) ? 0.0f : __ldg(tab + off);
}
HK_DEV float3 spectral_to_xyz(const DevTables& T, Spec L, float4 lambda, float4 pdf) {
    float3 s = f3(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float p = sp_get(pdf, i);
        if (p != 0.0f) {
            float l = sp_get(lambda, i);
            float3 cmf = f3(cie_lookup(T.cie_x, l), cie_lookup(T.cie_y, l), cie_lookup(T.cie_z, l));
            s = s + cmf * sp_get(L, i) / p;
        }
    }
    return s * 0.25f;
}
HK_DEV float3 xyz_to_linear_srgb(float3 c) {
    return f3(3.2404542f * c.x - 1.5371385f * c.y - 0.4985314f * c.z,
              -0.9692660f * c.x + 1.8760108f * c.y + 0.0415560f * c.z,
              0.0556434f * c.x - 0.2040259f * c.y + 1.0572252f * c.z);
}

// ---- filter ------------------------------------------------------------------------------------------------
struct DevFilter {
    int32_t type; float rx, ry; int32_t nx, ny;
    const float* __restrict__ func; const float* __restrict__ mcdf; const float* __restrict__ mfunc; const float* __restrict__ ccdf;
    float dmin_x, dmin_y, dmax_x, dmax_y, func_integral;
};
HK_DEV float sample_tent(float u, float r) {
    if (u < 0.5f) return -r + r * sqrtf(2.0f * u);
    return r * (1.0f - sqrtf(2.0f * (1.0f - u)));
}
HK_DEV int filter_interval(const float* __restrict__ cdf, float u, int n) {   // filter.jl:727-741, 1-based result
    int lo = 1, hi = n + 1;
    const int iters = 32 - __clz(n);      // >= ceil(log2(n)): the interval [lo, hi) is one entry wide by then, further rounds change nothing
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        int mid = (lo + hi) >> 1;
        bool c = __ldg(cdf + mid - 1) <= u;
        lo = c ? mid : lo; hi = c ? hi : mid;
    }
    return lo;
}
HK_DEV void filter_sample(const DevFilter& F, float2 u, float& px, float& py, float& w) {
    if (F.type == 1) { px = lerpf(-F.rx, F.rx, u.x); py = lerpf(-F.ry, F.ry, u.y); w = 1.0f; return; }
    if (F.type == 2) { px = sample_tent(u.x, F.rx); py = sample_tent(u.y, F.ry); w = 1.0f; return; }
    int o = clampi(filter_interval(F.mcdf, u.y, F.ny), 1, F.ny);
    float c0 = __ldg(F.mcdf + o - 1), c1 = __ldg(F.mcdf + o);
    float du = u.y - c0, df = c1 - c0;
    du = df > 0.0f ? du / df : 0.0f;
    float mf = __ldg(F.mfunc + o - 1);
    float pdf_y = F.func_integral > 0.0f ? mf / F.func_integral : 0.0f;
    py = lerpf(F.dmin_y, F.dmax_y, ((float)(o - 1) + du) / (float)F.ny);
    const float* cc = F.ccdf + (size_t)(o - 1) * (F.nx + 1);
    int ox = clampi(filter_interval(cc, u.x, F.nx), 1, F.nx);
    float d0 = __ldg(cc + ox - 1), d1 = __ldg(cc + ox);
    float dux = u.x - d0, dfx = d1 - d0;
    dux = dfx > 0.0f ? dux / dfx : 0.0f;
    float fv = __ldg(F.func + (size_t)(o - 1) * F.nx + (ox - 1));
    float pdf_x = mf > 0.0f ? fv / mf : 0.0f;
    px = lerpf(F.dmin_x, F.dmax_x, ((float)(ox - 1) + dux) / (float)F.nx);
    float pdf = pdf_x * pdf_y;
    w = pdf > 0.0f ? fv / pdf : 0.0f;
}

// ---- camera ------------------------------------------------------------------------------------------------
HK_DEV void camera_generate_ray(const HkCamera& C, float fx, float fy, float2 lens, float3& o_out, float3& d_out) {
    float3 pc = xf_point(C.raster_to_camera, f3(fx, fy, 0.0f));
    float3 o = f3(0, 0, 0);
    float3 d = norm3(pc);
    if (C.lens_radius > 0.0f) {
        float2 dl = concentric_sample_disk(lens);
        float t = -C.focal_distance / d.z;
        float3 pf = o + d * t;
        o = f3(C.lens_radius * dl.x, C.lens_radius * dl.y, 0.0f);
        d = norm3(pf - o);
    }
    o_out = xf_point(C.camera_to_world, o);
    d_out = norm3(xf_vec(C.camera_to_world, d));
}
