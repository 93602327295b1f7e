// ok_bsdf.h — ORACLE (test infrastructure, NOT product code).
// Spectral BSDF sample/eval per material, restating src/materials/spectral-eval.jl.
#pragma once
#include <vector>
#include "ok_spectral.h"

namespace ok {

struct BSDFSample {   // SpectralBSDFSample, spectral-eval.jl:18-27
    V3 wi; Spec f; float pdf; bool is_specular; float eta_scale;
    BSDFSample() : wi(0, 0, 1), f(), pdf(0.0f), is_specular(false), eta_scale(1.0f) {}
    BSDFSample(V3 w, Spec ff, float p, bool s, float e) : wi(w), f(ff), pdf(p), is_specular(s), eta_scale(e) {}
};
struct BSDFEval { Spec f; float pdf; BSDFEval() : f(), pdf(0.0f) {} BSDFEval(Spec ff, float p) : f(ff), pdf(p) {} };

struct TextureStore { std::vector<std::vector<float>> rgb; std::vector<int32_t> h, w; };
struct MatCtx {   // what the reference passes around as (table, textures) + the TextureFilterContext's uv
    const Tables* T;
    const HkSpectra* spectra;
    const TextureStore* textures = nullptr;
    V2 uv = V2(0.0f, 0.0f);
    uint32_t face_idx = 0; float bary[3] = {0.0f, 0.0f, 0.0f};      // TextureFilterContext.face_idx / bary (vertex colours)
};
// _sample_texture_bilinear, src/textures/texture-ref.jl:160-190, on an (h, w) column-major RGB image
inline void sample_texture_bilinear(const TextureStore& S, int32_t id, V2 uv, float* out) {
    const std::vector<float>& d = S.rgb[id - 1];
    const int h = S.h[id - 1], w = S.w[id - 1];
    const float ua0 = 1.0f - uv.y, ua1 = uv.x;                       // uv_adj = (1 - v, u)
    const float px = ua1 * (float)(w - 1) + 1.0f, py = ua0 * (float)(h - 1) + 1.0f;
    int x0 = (int)std::floor(px), y0 = (int)std::floor(py);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = clampi(x0, 1, w); x1 = clampi(x1, 1, w); y0 = clampi(y0, 1, h); y1 = clampi(y1, 1, h);
    const float fx = px - std::floor(px), fy = py - std::floor(py);
    auto at = [&](int y, int x, int c) { return d[3 * ((size_t)(x - 1) * h + (y - 1)) + c]; };
    for (int c = 0; c < 3; c++) {
        const float c0 = at(y0, x0, c) * (1.0f - fx) + at(y0, x1, c) * fx;
        const float c1 = at(y1, x0, c) * (1.0f - fx) + at(y1, x1, c) * fx;
        out[c] = c0 * (1.0f - fy) + c1 * fy;
    }
}
// eval_tex(textures, mat.Kd, tfc) for a MatteMaterial: the constant, or the bilinear texel at the hit's uv (texture-ref.jl:72-80)
inline void matte_kd_rgb(const MatCtx& C, const HkMaterial& m, float* kd) {
    if (m.tex[0] > 0 && C.textures && (m.flags & HK_MATFLAG_VERTEX_COLORS)) {
        // eval_tex(::VertexColorTexture, tfc), texture-ref.jl:240-245: data[1, fi] b1 + data[2, fi] b2 + data[3, fi] b3 on a (3, n_faces) table
        const std::vector<float>& d = C.textures->rgb[m.tex[0] - 1];
        const size_t base = 9 * (size_t)(C.face_idx - 1);
        for (int c = 0; c < 3; c++) kd[c] = d[base + c] * C.bary[0] + d[base + 3 + c] * C.bary[1] + d[base + 6 + c] * C.bary[2];
    }
    else if (m.tex[0] > 0 && C.textures) sample_texture_bilinear(*C.textures, m.tex[0], C.uv, kd);
    else { kd[0] = m.rgb0[0]; kd[1] = m.rgb0[1]; kd[2] = m.rgb0[2]; }
}

// src/reflection/bxdf.jl:67-90
inline float fresnel_dielectric(float cos_i, float eta) {
    cos_i = clampf(cos_i, -1.0f, 1.0f);
    if (cos_i < 0.0f) { eta = 1.0f / eta; cos_i = -cos_i; }
    float sin2_i = 1.0f - cos_i * cos_i;
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return 1.0f;
    float cos_t = std::sqrt(1.0f - sin2_t);
    float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// src/reflection/microfacet.jl:83-99
inline float roughness_to_alpha(float r) { return std::sqrt(r); }
inline float regularize_alpha(float a) { return a < 0.3f ? clampf(2.0f * a, 0.1f, 0.3f) : a; }

// local-frame trig helpers, spectral-eval.jl:3589-3647
inline float cos2_theta(V3 w) { return w.z * w.z; }
inline float abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline float sin2_theta(V3 w) { return std::max(0.0f, 1.0f - cos2_theta(w)); }
inline float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline float cos_phi(V3 w) { float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
inline float sin_phi(V3 w) { float s = sin_theta(w); return s == 0.0f ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
inline bool same_hemisphere(V3 a, V3 b) { return a.z * b.z > 0.0f; }
inline V3 face_forward(V3 v, V3 n) { return dot(v, n) < 0.0f ? -v : v; }

// spectral-eval.jl:3663-3739
inline float fr_complex(float cos_i, float eta, float k) {
    cos_i = clampf(cos_i, 0.0f, 1.0f);
    float sin2_i = 1.0f - cos_i * cos_i;
    float eta2 = eta * eta, k2 = k * k;
    float e_re = eta2 - k2, e_im = 2.0f * eta * k;
    float denom = e_re * e_re + e_im * e_im;
    float s2t_re = sin2_i * e_re / denom;
    float s2t_im = -sin2_i * e_im / denom;
    float c2t_re = 1.0f - s2t_re;
    float c2t_im = -s2t_im;
    float mag = std::sqrt(c2t_re * c2t_re + c2t_im * c2t_im);
    float ct_re = std::sqrt(0.5f * (mag + c2t_re));
    float ct_im = c2t_im / (2.0f * ct_re);
    if (ct_re == 0.0f) ct_im = std::sqrt(0.5f * mag);
    float eci_re = eta * cos_i, eci_im = k * cos_i;
    float np_re = eci_re - ct_re, np_im = eci_im - ct_im;
    float dp_re = eci_re + ct_re, dp_im = eci_im + ct_im;
    float dp_m2 = dp_re * dp_re + dp_im * dp_im;
    float rp_re = (np_re * dp_re + np_im * dp_im) / dp_m2;
    float rp_im = (np_im * dp_re - np_re * dp_im) / dp_m2;
    float ect_re = eta * ct_re - k * ct_im;
    float ect_im = eta * ct_im + k * ct_re;
    float ns_re = cos_i - ect_re, ns_im = -ect_im;
    float ds_re = cos_i + ect_re, ds_im = ect_im;
    float ds_m2 = ds_re * ds_re + ds_im * ds_im;
    float rs_re = (ns_re * ds_re + ns_im * ds_im) / ds_m2;
    float rs_im = (ns_im * ds_re - ns_re * ds_im) / ds_m2;
    float norm_parl = rp_re * rp_re + rp_im * rp_im;
    float norm_perp = rs_re * rs_re + rs_im * rs_im;
    return (norm_parl + norm_perp) * 0.5f;
}
inline Spec fr_complex_spectral(float c, const Spec& eta, const Spec& k) {
    return Spec(fr_complex(c, eta.v[0], k.v[0]), fr_complex(c, eta.v[1], k.v[1]), fr_complex(c, eta.v[2], k.v[2]), fr_complex(c, eta.v[3], k.v[3]));
}

// Trowbridge-Reitz, spectral-eval.jl:3765-3864
inline bool tr_effectively_smooth(float ax, float ay) { return std::max(ax, ay) < 1.0e-3f; }
inline float tr_d(V3 wm, float ax, float ay) {
    float t2 = tan2_theta(wm);
    if (std::isinf(t2)) return 0.0f;
    float c4 = cos2_theta(wm) * cos2_theta(wm);
    if (c4 < 1.0e-16f) return 0.0f;
    float a = cos_phi(wm) / ax, b = sin_phi(wm) / ay;
    float e = t2 * (a * a + b * b);
    float ope = 1.0f + e;
    return 1.0f / (PI_F * ax * ay * c4 * (ope * ope));
}
inline float tr_lambda(V3 w, float ax, float ay) {
    float t2 = tan2_theta(w);
    if (std::isinf(t2)) return 0.0f;
    float a = cos_phi(w) * ax, b = sin_phi(w) * ay;
    float alpha2 = a * a + b * b;
    return (std::sqrt(1.0f + alpha2 * t2) - 1.0f) * 0.5f;
}
inline float tr_g1(V3 w, float ax, float ay) { return 1.0f / (1.0f + tr_lambda(w, ax, ay)); }
inline float tr_g(V3 wo, V3 wi, float ax, float ay) { return 1.0f / (1.0f + tr_lambda(wo, ax, ay) + tr_lambda(wi, ax, ay)); }
inline float tr_pdf(V3 w, V3 wm, float ax, float ay) {
    return tr_g1(w, ax, ay) / abs_cos_theta(w) * tr_d(wm, ax, ay) * std::fabs(dot(w, wm));
}
inline V3 tr_sample_wm(V3 w, V2 u, float ax, float ay) {
    V3 wh = normalize(V3(ax * w.x, ay * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    V3 t1 = wh.z < 0.99999f ? normalize(cross(V3(0, 0, 1), wh)) : V3(1, 0, 0);
    V3 t2 = cross(wh, t1);
    float r = std::sqrt(u.x);
    float phi = 2.0f * PI_F * u.y;
    float px = r * dm_cosf(phi);
    float py = r * dm_sinf(phi);
    float h = std::sqrt(1.0f - px * px);
    py = lerpf(h, py, 0.5f * (1.0f + wh.z));
    float pz = std::sqrt(std::max(0.0f, 1.0f - px * px - py * py));
    V3 nh = px * t1 + py * t2 + pz * wh;
    return normalize(V3(ax * nh.x, ay * nh.y, std::max(1.0e-6f, nh.z)));
}

// clamp(::RGBSpectrum) to [0,1] (src/spectrum.jl)
inline void clamp_rgb01(const float* in, float* out) { for (int i = 0; i < 3; i++) out[i] = clampf(in[i], 0.0f, 1.0f); }

// spectral-eval.jl:206-210  eval_ior_spectral
inline Spec eval_ior_spectral(const MatCtx& C, const HkMaterial& m, int which, const Wavelengths& l) {
    if ((m.flags & HK_MATFLAG_SPECTRAL_ETA_K) && m.spec[which] > 0) {
        uint32_t id = (uint32_t)m.spec[which];
        uint32_t a = C.spectra->offsets[id - 1], b = C.spectra->offsets[id];
        Spec r;
        for (int i = 0; i < 4; i++) r.v[i] = pls_sample(C.spectra->lambdas + a, C.spectra->values + a, (int)(b - a), l.lambda[i]);
        return r;
    }
    return uplift_rgb_unbounded(*C.T, which == 0 ? m.rgb0 : m.rgb1, l);
}

// ---------------------------------------------------------------------------------------------
// Matte  spectral-eval.jl:42-101 (sample), 371-397 (eval)
// ---------------------------------------------------------------------------------------------
inline BSDFSample sample_matte(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2 u, float /*rng*/, bool /*reg*/) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    float kd_raw[3], kd[3]; matte_kd_rgb(C, m, kd_raw); clamp_rgb01(kd_raw, kd);
    float sigma = m.f[0];
    Spec kd_s = uplift_rgb(*C.T, kd, l);
    V3 t, b; coordinate_system(n, t, b);
    V3 lw = cosine_sample_hemisphere(u);
    float cos_theta = lw.z;
    if (cos_theta < 1.0e-6f) return BSDFSample();
    if (wo_dot_n < 0.0f) lw = V3(lw.x, lw.y, -lw.z);
    V3 wi = normalize(local_to_world(lw, n, t, b));
    Spec f;
    if (sigma > 0.0f) {
        float rf = 1.0f - 0.5f * sigma / (sigma + 0.33f);
        f = kd_s * (rf / PI_F);
    } else {
        f = kd_s * (1.0f / PI_F);
    }
    return BSDFSample(wi, f, cos_theta / PI_F, false, 1.0f);
}
inline BSDFEval eval_matte(const MatCtx& C, const HkMaterial& m, V3 wo, V3 wi, V3 n, const Wavelengths& l) {
    float ci = dot(wi, n), co = dot(wo, n);
    if (ci * co < 0.0f) return BSDFEval();
    float c = std::fabs(ci);
    if (c < 1.0e-6f) return BSDFEval();
    float kd_raw[3], kd[3]; matte_kd_rgb(C, m, kd_raw); clamp_rgb01(kd_raw, kd);
    Spec kd_s = uplift_rgb(*C.T, kd, l);
    return BSDFEval(kd_s / PI_F, c / PI_F);
}

// ---------------------------------------------------------------------------------------------
// Mirror  spectral-eval.jl:108-132
// ---------------------------------------------------------------------------------------------
inline BSDFSample sample_mirror(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2, float, bool) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    Spec kr = uplift_rgb(*C.T, m.rgb0, l);
    V3 no = wo_dot_n < 0.0f ? -n : n;
    return BSDFSample(reflect(wo, no), kr, 1.0f, true, 1.0f);
}

// ---------------------------------------------------------------------------------------------
// Glass  spectral-eval.jl:140-198
// ---------------------------------------------------------------------------------------------
inline BSDFSample sample_glass(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2, float rng, bool) {
    float ior = m.f[0];
    if (ior == 0.0f) ior = 1.0f;
    Spec kr = uplift_rgb(*C.T, m.rgb0, l);
    Spec kt = uplift_rgb(*C.T, m.rgb1, l);
    float cos_o = dot(wo, n);
    bool entering = cos_o > 0.0f;
    V3 no = entering ? n : -n;
    cos_o = std::fabs(cos_o);
    float eta = entering ? ior : (1.0f / ior);
    float F = fresnel_dielectric(cos_o, eta);
    if (rng < F) return BSDFSample(reflect(wo, no), kr, 1.0f, true, 1.0f);
    float sin2_i = std::max(0.0f, 1.0f - cos_o * cos_o);
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return BSDFSample(reflect(wo, no), kr, 1.0f, true, 1.0f);
    float cos_t = std::sqrt(1.0f - sin2_t);
    V3 wi = normalize(-wo / eta + (cos_o / eta - cos_t) * no);
    return BSDFSample(wi, kt, 1.0f, true, 1.0f / (eta * eta));
}

// ---------------------------------------------------------------------------------------------
// Conductor  spectral-eval.jl:223-318 (sample), 421-488 (eval)
// ---------------------------------------------------------------------------------------------
inline BSDFSample sample_conductor(const MatCtx& C, const HkMaterial& m, V3 wo_w, V3 n, const Wavelengths& l, V2 u, float, bool regularize) {
    V3 t, b; coordinate_system(n, t, b);
    V3 wo = world_to_local(wo_w, n, t, b);
    if (wo.z == 0.0f) return BSDFSample();
    float rough = m.f[0];
    float ax = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) ? roughness_to_alpha(rough) : rough;
    float ay = ax;
    if (regularize) { ax = regularize_alpha(ax); ay = regularize_alpha(ay); }
    if (!tr_effectively_smooth(ax, ay)) { ax = std::max(ax, 1.0e-4f); ay = std::max(ay, 1.0e-4f); }
    Spec eta = eval_ior_spectral(C, m, 0, l);
    Spec k = eval_ior_spectral(C, m, 1, l);
    if (tr_effectively_smooth(ax, ay)) {
        V3 wi(-wo.x, -wo.y, wo.z);
        float ci = abs_cos_theta(wi);
        Spec F = fr_complex_spectral(ci, eta, k);
        return BSDFSample(local_to_world(wi, n, t, b), F / ci, 1.0f, true, 1.0f);
    }
    V3 wm = tr_sample_wm(wo, u, ax, ay);
    V3 wi = -wo + 2.0f * dot(wo, wm) * wm;
    if (!same_hemisphere(wo, wi)) return BSDFSample();
    float pdf = tr_pdf(wo, wm, ax, ay) / (4.0f * std::fabs(dot(wo, wm)));
    float co = abs_cos_theta(wo), ci = abs_cos_theta(wi);
    if (ci == 0.0f || co == 0.0f) return BSDFSample();
    Spec F = fr_complex_spectral(std::fabs(dot(wo, wm)), eta, k);
    float D = tr_d(wm, ax, ay), G = tr_g(wo, wi, ax, ay);
    Spec f = D * F * G / (4.0f * ci * co);
    return BSDFSample(local_to_world(wi, n, t, b), f, pdf, false, 1.0f);
}
inline BSDFEval eval_conductor(const MatCtx& C, const HkMaterial& m, V3 wo_w, V3 wi_w, V3 n, const Wavelengths& l) {
    V3 t, b; coordinate_system(n, t, b);
    V3 wo = world_to_local(wo_w, n, t, b), wi = world_to_local(wi_w, n, t, b);
    if (!same_hemisphere(wo, wi)) return BSDFEval();
    float rough = m.f[0];
    float ax = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) ? roughness_to_alpha(rough) : rough;
    float ay = ax;
    if (!tr_effectively_smooth(ax, ay)) { ax = std::max(ax, 1.0e-4f); ay = std::max(ay, 1.0e-4f); }
    if (tr_effectively_smooth(ax, ay)) return BSDFEval();
    float co = abs_cos_theta(wo), ci = abs_cos_theta(wi);
    if (ci == 0.0f || co == 0.0f) return BSDFEval();
    V3 wm = wi + wo;
    if (dot(wm, wm) == 0.0f) return BSDFEval();
    wm = normalize(wm);
    Spec eta = eval_ior_spectral(C, m, 0, l);
    Spec k = eval_ior_spectral(C, m, 1, l);
    Spec F = fr_complex_spectral(std::fabs(dot(wo, wm)), eta, k);
    float D = tr_d(wm, ax, ay), G = tr_g(wo, wi, ax, ay);
    Spec f = D * F * G / (4.0f * ci * co);
    V3 wmp = face_forward(wm, V3(0, 0, 1));
    float pdf = tr_pdf(wo, wmp, ax, ay) / (4.0f * std::fabs(dot(wo, wmp)));
    return BSDFEval(f, pdf);
}

}  // namespace ok
#include "ok_bsdf_layered.h"
#include "ok_bsdf_coated_conductor.h"
#include "ok_bsdf_coated_difftrans.h"
namespace ok {

// ---------------------------------------------------------------------------------------------
// Dispatch  src/integrators/physical-wavefront/material-dispatch.jl:23-53
// ---------------------------------------------------------------------------------------------
// eval_tex(textures, mat.<param>, tfc) at every parameter read of spectral-eval.jl: a per-hit copy of the material with its textured
// RGB / scalar parameters replaced by the bilinear texel at the hit's uv (HkMaterial.tex / ftex); Matte.Kd keeps its own path above
inline HkMaterial resolve_material_textures(const MatCtx& C, const HkMaterial& g) {
    HkMaterial out = g;
    if (!C.textures) return out;
    float rgb[3];
    for (int k = 0; k < 3; k++) {
        if (g.tex[k] <= 0 || (k == 0 && g.type == HK_MAT_MATTE)) continue;
        sample_texture_bilinear(*C.textures, g.tex[k], C.uv, rgb);
        float* dst = k == 0 ? out.rgb0 : (k == 1 ? out.rgb1 : out.rgb2);
        dst[0] = rgb[0]; dst[1] = rgb[1]; dst[2] = rgb[2];
    }
    for (int k = 0; k < 8; k++) {
        if (g.ftex[k] <= 0) continue;
        sample_texture_bilinear(*C.textures, g.ftex[k], C.uv, rgb);
        out.f[k] = rgb[0];
    }
    return out;
}
inline BSDFSample sample_material(const MatCtx& C, const HkMaterial& m_in, V3 wo, V3 ns, const Wavelengths& l, V2 u, float rng, bool regularize) {
    const HkMaterial m = resolve_material_textures(C, m_in);
    switch (m.type) {
        case HK_MAT_MATTE: return sample_matte(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_MIRROR: return sample_mirror(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_GLASS: return sample_glass(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_CONDUCTOR: return sample_conductor(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_COATED_DIFFUSE: return sample_coated_diffuse(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_THIN_DIELECTRIC: return sample_thin_dielectric(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_DIFFUSE_TRANSMISSION: return sample_diffuse_transmission(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_COATED_CONDUCTOR: return sample_coated_conductor(C, m, wo, ns, l, u, rng, regularize);
        case HK_MAT_COATED_DIFFUSE_TRANSMISSION: return sample_coated_diffuse_transmission(C, m, wo, ns, l, u, rng, regularize);
    }
    return BSDFSample();
}
inline BSDFEval eval_material(const MatCtx& C, const HkMaterial& m_in, V3 wo, V3 wi, V3 ns, const Wavelengths& l) {
    const HkMaterial m = resolve_material_textures(C, m_in);
    switch (m.type) {
        case HK_MAT_MATTE: return eval_matte(C, m, wo, wi, ns, l);
        case HK_MAT_CONDUCTOR: return eval_conductor(C, m, wo, wi, ns, l);
        case HK_MAT_COATED_DIFFUSE: return eval_coated_diffuse(C, m, wo, wi, ns, l);
        case HK_MAT_DIFFUSE_TRANSMISSION: return eval_diffuse_transmission(C, m, wo, wi, ns, l);
        case HK_MAT_COATED_CONDUCTOR: return eval_coated_conductor(C, m, wo, wi, ns, l);
        case HK_MAT_COATED_DIFFUSE_TRANSMISSION: return eval_coated_diffuse_transmission(C, m, wo, wi, ns, l);
        default: return BSDFEval();   // Mirror / Glass / ThinDielectric: specular, eval == 0
    }
}

}  // namespace ok
