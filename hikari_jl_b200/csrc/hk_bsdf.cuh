// hk_bsdf.cuh — spectral BSDF sample / eval device functions, one pair per material type so each
// per-material queue kernel instantiates only its own code (no megakernel type switch).
// Reference: src/materials/spectral-eval.jl (line ranges at each function), src/reflection/bxdf.jl:67-90,
// src/reflection/microfacet.jl:83-99.
#pragma once
#include "hk_spectral.cuh"

struct BsdfSample { float3 wi; Spec f; float pdf; bool specular; float eta_scale; };
struct BsdfEval { Spec f; float pdf; };
HK_DEV BsdfSample bsdf_none() { BsdfSample s; s.wi = f3(0, 0, 1); s.f = sp(0.0f); s.pdf = 0.0f; s.specular = false; s.eta_scale = 1.0f; return s; }
HK_DEV BsdfSample bsdf_make(float3 wi, Spec f, float pdf, bool spec, float eta) { BsdfSample s; s.wi = wi; s.f = f; s.pdf = pdf; s.specular = spec; s.eta_scale = eta; return s; }
HK_DEV BsdfEval eval_none() { BsdfEval e; e.f = sp(0.0f); e.pdf = 0.0f; return e; }
HK_DEV BsdfEval eval_make(Spec f, float pdf) { BsdfEval e; e.f = f; e.pdf = pdf; return e; }

// local: the material being shaded is a per-hit copy whose textured parameters were resolved at the hit (resolve_material_textures):
// the per-material uplift cache does not apply to it
struct MatCtx { DevTables T; const float* __restrict__ spec_lambdas; const float* __restrict__ spec_values; const uint32_t* __restrict__ spec_offsets; bool local; };

HK_DEV float fresnel_dielectric(float ci, float eta) {   // bxdf.jl:67-90
    ci = clampf(ci, -1.0f, 1.0f);
    if (ci < 0.0f) { eta = 1.0f / eta; ci = -ci; }
    float s2i = 1.0f - ci * ci;
    float s2t = s2i / (eta * eta);
    if (s2t >= 1.0f) return 1.0f;
    float ct = sqrtf(1.0f - s2t);
    float rp = (eta * ci - ct) / (eta * ci + ct);
    float rs = (ci - eta * ct) / (ci + eta * ct);
    return 0.5f * (rp * rp + rs * rs);
}
HK_DEV float regularize_alpha(float a) { return a < 0.3f ? clampf(2.0f * a, 0.1f, 0.3f) : a; }

// local-frame trigonometry (spectral-eval.jl:3589-3647)
HK_DEV float cos2_t(float3 w) { return w.z * w.z; }
HK_DEV float sin2_t(float3 w) { return fmaxf(0.0f, 1.0f - cos2_t(w)); }
HK_DEV float tan2_t(float3 w) { return sin2_t(w) / cos2_t(w); }
HK_DEV float cos_phi(float3 w) { float s = sqrtf(sin2_t(w)); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
HK_DEV float sin_phi(float3 w) { float s = sqrtf(sin2_t(w)); return s == 0.0f ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
HK_DEV bool same_hemi(float3 a, float3 b) { return a.z * b.z > 0.0f; }

HK_DEV float fr_complex(float ci, float eta, float k) {   // :3663-3739
    ci = clampf(ci, 0.0f, 1.0f);
    float s2i = 1.0f - ci * ci;
    float e_re = eta * eta - k * k, e_im = 2.0f * eta * k;
    float den = e_re * e_re + e_im * e_im;
    float s_re = s2i * e_re / den, s_im = -s2i * e_im / den;
    float c_re = 1.0f - s_re, c_im = -s_im;
    float mag = sqrtf(c_re * c_re + c_im * c_im);
    float t_re = sqrtf(0.5f * (mag + c_re));
    float t_im = c_im / (2.0f * t_re);
    if (t_re == 0.0f) t_im = sqrtf(0.5f * mag);
    float a_re = eta * ci, a_im = k * ci;
    float n_re = a_re - t_re, n_im = a_im - t_im, d_re = a_re + t_re, d_im = a_im + t_im;
    float dm = d_re * d_re + d_im * d_im;
    float p_re = (n_re * d_re + n_im * d_im) / dm, p_im = (n_im * d_re - n_re * d_im) / dm;
    float b_re = eta * t_re - k * t_im, b_im = eta * t_im + k * t_re;
    float m_re = ci - b_re, m_im = -b_im, q_re = ci + b_re, q_im = b_im;
    float qm = q_re * q_re + q_im * q_im;
    float s2_re = (m_re * q_re + m_im * q_im) / qm, s2_im = (m_im * q_re - m_re * q_im) / qm;
    return ((p_re * p_re + p_im * p_im) + (s2_re * s2_re + s2_im * s2_im)) * 0.5f;
}
HK_NI_BSDF Spec fr_complex4(float c, Spec eta, Spec k) { return sp4(fr_complex(c, eta.x, k.x), fr_complex(c, eta.y, k.y), fr_complex(c, eta.z, k.z), fr_complex(c, eta.w, k.w)); }

// Trowbridge-Reitz (:3765-3864)
HK_DEV bool tr_smooth(float ax, float ay) { return fmaxf(ax, ay) < 1.0e-3f; }
HK_NI_BSDF float tr_d(float3 wm, float ax, float ay) {
    float t2 = tan2_t(wm);
    if (isinf(t2)) return 0.0f;
    float c4 = cos2_t(wm) * cos2_t(wm);
    if (c4 < 1.0e-16f) return 0.0f;
    float a = cos_phi(wm) / ax, b = sin_phi(wm) / ay;
    float e = t2 * (a * a + b * b);
    float q = 1.0f + e;
    return 1.0f / (HK_PI * ax * ay * c4 * (q * q));
}
HK_NI_BSDF float tr_lambda(float3 w, float ax, float ay) {
    float t2 = tan2_t(w);
    if (isinf(t2)) return 0.0f;
    float a = cos_phi(w) * ax, b = sin_phi(w) * ay;
    return (sqrtf(1.0f + (a * a + b * b) * t2) - 1.0f) * 0.5f;
}
HK_DEV float tr_g1(float3 w, float ax, float ay) { return 1.0f / (1.0f + tr_lambda(w, ax, ay)); }
HK_DEV float tr_g(float3 wo, float3 wi, float ax, float ay) { return 1.0f / (1.0f + tr_lambda(wo, ax, ay) + tr_lambda(wi, ax, ay)); }
HK_DEV float tr_pdf(float3 w, float3 wm, float ax, float ay) { return tr_g1(w, ax, ay) / fabsf(w.z) * tr_d(wm, ax, ay) * fabsf(dot3(w, wm)); }
HK_NI_BSDF float3 tr_sample_wm(float3 w, float2 u, float ax, float ay) {
    float3 wh = norm3(f3(ax * w.x, ay * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    float3 t1 = wh.z < 0.99999f ? norm3(cross3(f3(0, 0, 1), wh)) : f3(1, 0, 0);
    float3 t2 = cross3(wh, t1);
    float r = sqrtf(u.x), phi = 2.0f * HK_PI * u.y;
    float px = r * dm_cosf(phi), py = r * dm_sinf(phi);
    float h = sqrtf(1.0f - px * px);
    py = lerpf(h, py, 0.5f * (1.0f + wh.z));
    float pz = sqrtf(fmaxf(0.0f, 1.0f - px * px - py * py));
    float3 nh = px * t1 + py * t2 + pz * wh;
    return norm3(f3(ax * nh.x, ay * nh.y, fmaxf(1.0e-6f, nh.z)));
}
// Uplift cache slots per material type (DevTables::mat_pre): the constant RGB a BSDF uplifts at every call.
//   Matte / Mirror: A = rgb0.  Glass / CoatedDiffuse: A = rgb0, B = rgb1.  Conductor (RGB eta / k): A, B unbounded.
//   DiffuseTransmission: A / B = clamp(rgb0 * scale) / clamp(rgb1 * scale).
//   CoatedConductor: A = eta (unbounded) or the reflectance clamped to [0, 0.9999]; B = k (unbounded).
//   CoatedDiffuseTransmission: A = reflectance, B = albedo (as CoatedDiffuse); the transmittance (rgb2) is uplifted per call.
HK_DEV bool mat_pre_is_unbounded(const HkMaterial& m, int which) {
    return m.type == HK_MAT_CONDUCTOR || (m.type == HK_MAT_COATED_CONDUCTOR && (which == 1 || (m.flags & HK_MATFLAG_USE_ETA_K)));
}
HK_DEV float4 mat_pre_compute(const DevTables& T, const HkMaterial& m, int which) {
    const float* c = which == 0 ? m.rgb0 : m.rgb1;
    if (mat_pre_is_unbounded(m, which)) return make_pre_unbounded(T, c[0], c[1], c[2]);
    if (m.type == HK_MAT_COATED_CONDUCTOR) return make_pre_bounded(T, clampf(c[0], 0.0f, 0.9999f), clampf(c[1], 0.0f, 0.9999f), clampf(c[2], 0.0f, 0.9999f));
    if (m.type == HK_MAT_DIFFUSE_TRANSMISSION) { const float s = m.f[0]; return make_pre_bounded(T, clampf(c[0] * s, 0.0f, 1.0f), clampf(c[1] * s, 0.0f, 1.0f), clampf(c[2] * s, 0.0f, 1.0f)); }
    return make_pre_bounded(T, c[0], c[1], c[2]);      // (rgb_to_spectrum clamps to [0,1] itself)
}
HK_DEV Spec mat_spec(const MatCtx& C, const HkMaterial& m, int which, float4 lam) {
    const float4 q = (C.T.mat_pre && !C.local) ? __ldg(C.T.mat_pre + 2 * (&m - (const HkMaterial*)C.T.mat_base) + which) : mat_pre_compute(C.T, m, which);
    return mat_pre_is_unbounded(m, which) ? pre_unbounded(q, lam) : pre_bounded(q, lam);
}
HK_DEV Spec ior_spectrum(const MatCtx& C, const HkMaterial& m, int which, float4 lambda) {   // :206-210
    if ((m.flags & HK_MATFLAG_SPECTRAL_ETA_K) && m.spec[which] > 0) {
        uint32_t a = __ldg(C.spec_offsets + m.spec[which] - 1), b = __ldg(C.spec_offsets + m.spec[which]);
        const float* l = C.spec_lambdas + a; const float* v = C.spec_values + a; int n = (int)(b - a);
        return sp4(pls_sample(l, v, n, lambda.x), pls_sample(l, v, n, lambda.y), pls_sample(l, v, n, lambda.z), pls_sample(l, v, n, lambda.w));
    }
    return mat_spec(C, m, which, lambda);
}

// ---- Matte :42-101 / :371-397 -----------------------------------------------------------------------------
HK_DEV BsdfSample sample_matte(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam, float2 u) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    Spec kd = mat_spec(C, m, 0, lam);
    Frame fr = make_frame(n);
    float3 lw = cosine_sample_hemisphere(u);
    float ct = lw.z;
    if (ct < 1.0e-6f) return bsdf_none();
    if (wn < 0.0f) lw.z = -lw.z;
    float3 wi = norm3(to_world(fr, lw));
    float sigma = m.f[0];
    Spec f = sigma > 0.0f ? kd * ((1.0f - 0.5f * sigma / (sigma + 0.33f)) / HK_PI) : kd * (1.0f / HK_PI);
    return bsdf_make(wi, f, ct / HK_PI, false, 1.0f);
}
HK_DEV BsdfEval eval_matte(const MatCtx& C, const HkMaterial& m, float3 wo, float3 wi, float3 n, float4 lam) {
    float ci = dot3(wi, n), co = dot3(wo, n);
    if (ci * co < 0.0f) return eval_none();
    float c = fabsf(ci);
    if (c < 1.0e-6f) return eval_none();
    Spec kd = mat_spec(C, m, 0, lam);
    return eval_make(kd / HK_PI, c / HK_PI);
}
// Matte with a textured Kd: the same two functions with the uplifted texel passed in (k_shade<HK_SHADE_MATTE_TEX>)
HK_DEV BsdfSample sample_matte_kd(Spec kd, float sigma, float3 wo, float3 n, float2 u) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    Frame fr = make_frame(n);
    float3 lw = cosine_sample_hemisphere(u);
    float ct = lw.z;
    if (ct < 1.0e-6f) return bsdf_none();
    if (wn < 0.0f) lw.z = -lw.z;
    float3 wi = norm3(to_world(fr, lw));
    Spec f = sigma > 0.0f ? kd * ((1.0f - 0.5f * sigma / (sigma + 0.33f)) / HK_PI) : kd * (1.0f / HK_PI);
    return bsdf_make(wi, f, ct / HK_PI, false, 1.0f);
}
HK_DEV BsdfEval eval_matte_kd(Spec kd, float3 wo, float3 wi, float3 n) {
    float ci = dot3(wi, n), co = dot3(wo, n);
    if (ci * co < 0.0f) return eval_none();
    float c = fabsf(ci);
    if (c < 1.0e-6f) return eval_none();
    return eval_make(kd / HK_PI, c / HK_PI);
}
// ---- Mirror :108-132 ---------------------------------------------------------------------------------------
HK_DEV BsdfSample sample_mirror(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    Spec kr = mat_spec(C, m, 0, lam);
    return bsdf_make(reflect3(wo, wn < 0.0f ? -n : n), kr, 1.0f, true, 1.0f);
}
// ---- Glass :140-198 ----------------------------------------------------------------------------------------
HK_DEV BsdfSample sample_glass(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam, float uc) {
    float ior = m.f[0] == 0.0f ? 1.0f : m.f[0];
    Spec kr = mat_spec(C, m, 0, lam);
    Spec kt = mat_spec(C, m, 1, lam);
    float co = dot3(wo, n);
    bool entering = co > 0.0f;
    float3 no = entering ? n : -n;
    co = fabsf(co);
    float eta = entering ? ior : 1.0f / ior;
    float F = fresnel_dielectric(co, eta);
    if (uc < F) return bsdf_make(reflect3(wo, no), kr, 1.0f, true, 1.0f);
    float s2i = fmaxf(0.0f, 1.0f - co * co);
    float s2t = s2i / (eta * eta);
    if (s2t >= 1.0f) return bsdf_make(reflect3(wo, no), kr, 1.0f, true, 1.0f);
    float ct = sqrtf(1.0f - s2t);
    float3 wi = norm3(-wo / eta + (co / eta - ct) * no);
    return bsdf_make(wi, kt, 1.0f, true, 1.0f / (eta * eta));
}
// ---- Conductor :223-318 / :421-488 ----------------------------------------------------------------------------
HK_DEV BsdfSample sample_conductor(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 n, float4 lam, float2 u, bool regularize) {
    Frame fr = make_frame(n);
    float3 wo = to_local(fr, wo_w);
    if (wo.z == 0.0f) return bsdf_none();
    float ax = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) ? sqrtf(m.f[0]) : m.f[0];
    float ay = ax;
    if (regularize) { ax = regularize_alpha(ax); ay = regularize_alpha(ay); }
    if (!tr_smooth(ax, ay)) { ax = fmaxf(ax, 1.0e-4f); ay = fmaxf(ay, 1.0e-4f); }
    Spec eta = ior_spectrum(C, m, 0, lam), k = ior_spectrum(C, m, 1, lam);
    if (tr_smooth(ax, ay)) {
        float3 wi = f3(-wo.x, -wo.y, wo.z);
        float ci = fabsf(wi.z);
        return bsdf_make(to_world(fr, wi), fr_complex4(ci, eta, k) / ci, 1.0f, true, 1.0f);
    }
    float3 wm = tr_sample_wm(wo, u, ax, ay);
    float3 wi = -wo + 2.0f * dot3(wo, wm) * wm;
    if (!same_hemi(wo, wi)) return bsdf_none();
    float pdf = tr_pdf(wo, wm, ax, ay) / (4.0f * fabsf(dot3(wo, wm)));
    float co = fabsf(wo.z), ci = fabsf(wi.z);
    if (ci == 0.0f || co == 0.0f) return bsdf_none();
    Spec F = fr_complex4(fabsf(dot3(wo, wm)), eta, k);
    Spec f = tr_d(wm, ax, ay) * F * tr_g(wo, wi, ax, ay) / (4.0f * ci * co);
    return bsdf_make(to_world(fr, wi), f, pdf, false, 1.0f);
}
HK_DEV BsdfEval eval_conductor(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 wi_w, float3 n, float4 lam) {
    Frame fr = make_frame(n);
    float3 wo = to_local(fr, wo_w), wi = to_local(fr, wi_w);
    if (!same_hemi(wo, wi)) return eval_none();
    float ax = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) ? sqrtf(m.f[0]) : m.f[0];
    float ay = ax;
    if (!tr_smooth(ax, ay)) { ax = fmaxf(ax, 1.0e-4f); ay = fmaxf(ay, 1.0e-4f); }
    if (tr_smooth(ax, ay)) return eval_none();
    float co = fabsf(wo.z), ci = fabsf(wi.z);
    if (ci == 0.0f || co == 0.0f) return eval_none();
    float3 wm = wi + wo;
    if (dot3(wm, wm) == 0.0f) return eval_none();
    wm = norm3(wm);
    Spec eta = ior_spectrum(C, m, 0, lam), k = ior_spectrum(C, m, 1, lam);
    Spec F = fr_complex4(fabsf(dot3(wo, wm)), eta, k);
    Spec f = tr_d(wm, ax, ay) * F * tr_g(wo, wi, ax, ay) / (4.0f * ci * co);
    float3 wp = wm.z < 0.0f ? -wm : wm;
    return eval_make(f, tr_pdf(wo, wp, ax, ay) / (4.0f * fabsf(dot3(wo, wp))));
}
// ---- ThinDielectric :1975-2037 ----------------------------------------------------------------------------------
HK_DEV BsdfSample sample_thin_dielectric(const HkMaterial& m, float3 wo, float3 n, float uc) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    Frame fr = make_frame(n);
    float3 wl = f3(dot3(wo, fr.t), dot3(wo, fr.b), wn);
    float co = fabsf(wl.z);
    float R0 = fresnel_dielectric(co, m.f[0]), T0 = 1.0f - R0;
    float R = R0;
    if (R0 < 1.0f) R = R0 + T0 * T0 * R0 / (1.0f - R0 * R0);
    float T = 1.0f - R;
    if (R + T < 1.0e-10f) return bsdf_none();
    float pr = R / (R + T);
    if (uc < pr) {
        float3 il = f3(-wl.x, -wl.y, wl.z);
        return bsdf_make(norm3(to_world(fr, il)), sp(R / fabsf(il.z)), pr, true, 1.0f);
    }
    return bsdf_make(-wo, sp(T / co), 1.0f - pr, true, 1.0f);
}
// ---- DiffuseTransmission :2083-2218 -----------------------------------------------------------------------------
HK_DEV void difftrans_rgb(const HkMaterial& m, float* r, float* t) {
    float s = m.f[0];
#pragma unroll
    for (int i = 0; i < 3; i++) { r[i] = clampf(m.rgb0[i] * s, 0.0f, 1.0f); t[i] = clampf(m.rgb1[i] * s, 0.0f, 1.0f); }
}
HK_DEV BsdfSample sample_diffuse_transmission(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam, float2 u, float uc) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    float r[3], t[3]; difftrans_rgb(m, r, t);
    Spec rs = mat_spec(C, m, 0, lam), ts = mat_spec(C, m, 1, lam);
    float pr = fmaxf(fmaxf(r[0], r[1]), r[2]), pt = fmaxf(fmaxf(t[0], t[1]), t[2]);
    if (pr + pt < 1.0e-10f) return bsdf_none();
    Frame fr = make_frame(n);
    float prob = pr / (pr + pt);
    float3 lw = cosine_sample_hemisphere(u);
    bool refl = uc < prob;
    if (refl ? (wn < 0.0f) : (wn > 0.0f)) lw.z = -lw.z;
    float c = fabsf(lw.z);
    if (c < 1.0e-6f) return bsdf_none();
    float3 wi = norm3(to_world(fr, lw));
    if (refl) return bsdf_make(wi, rs * (1.0f / HK_PI), prob * c / HK_PI, false, 1.0f);
    return bsdf_make(wi, ts * (1.0f / HK_PI), (1.0f - prob) * c / HK_PI, false, 1.0f);
}
HK_DEV BsdfEval eval_diffuse_transmission(const MatCtx& C, const HkMaterial& m, float3 wo, float3 wi, float3 n, float4 lam) {
    float ci = dot3(wi, n), co = dot3(wo, n);
    float ac = fabsf(ci);
    if (ac < 1.0e-6f) return eval_none();
    float r[3], t[3]; difftrans_rgb(m, r, t);
    Spec rs = mat_spec(C, m, 0, lam), ts = mat_spec(C, m, 1, lam);
    float pr = fmaxf(fmaxf(r[0], r[1]), r[2]), pt = fmaxf(fmaxf(t[0], t[1]), t[2]);
    if (pr + pt < 1.0e-10f) return eval_none();
    if (ci * co > 0.0f) return eval_make(rs * (1.0f / HK_PI), pr / (pr + pt) * ac / HK_PI);
    return eval_make(ts * (1.0f / HK_PI), pt / (pr + pt) * ac / HK_PI);
}

#include "hk_bsdf_layered.cuh"
#include "hk_bsdf_coated_conductor.cuh"
#include "hk_bsdf_coated_difftrans.cuh"

// ---- per-type dispatch used by the per-material-queue kernels (TYPE is a compile-time constant) ------------------
template <int TYPE>
HK_DEV BsdfSample sample_bsdf(const MatCtx& C, const HkMaterial& m, float3 wo, float3 ns, float4 lam, float2 u, float uc, bool regularize) {
    if (TYPE == HK_MAT_MATTE) return sample_matte(C, m, wo, ns, lam, u);
    if (TYPE == HK_MAT_MIRROR) return sample_mirror(C, m, wo, ns, lam);
    if (TYPE == HK_MAT_GLASS) return sample_glass(C, m, wo, ns, lam, uc);
    if (TYPE == HK_MAT_CONDUCTOR) return sample_conductor(C, m, wo, ns, lam, u, regularize);
    if (TYPE == HK_MAT_COATED_DIFFUSE) return sample_coated_diffuse(C, m, wo, ns, lam, u, uc, regularize);
    if (TYPE == HK_MAT_THIN_DIELECTRIC) return sample_thin_dielectric(m, wo, ns, uc);
    if (TYPE == HK_MAT_DIFFUSE_TRANSMISSION) return sample_diffuse_transmission(C, m, wo, ns, lam, u, uc);
    if (TYPE == HK_MAT_COATED_CONDUCTOR) return sample_coated_conductor(C, m, wo, ns, lam, u, uc, regularize);
    if (TYPE == HK_MAT_COATED_DIFFUSE_TRANSMISSION) return sample_coated_difftrans(C, m, wo, ns, lam, u, uc, regularize);
    return bsdf_none();
}
template <int TYPE>
HK_DEV BsdfEval eval_bsdf(const MatCtx& C, const HkMaterial& m, float3 wo, float3 wi, float3 ns, float4 lam) {
    if (TYPE == HK_MAT_MATTE) return eval_matte(C, m, wo, wi, ns, lam);
    if (TYPE == HK_MAT_CONDUCTOR) return eval_conductor(C, m, wo, wi, ns, lam);
    if (TYPE == HK_MAT_COATED_DIFFUSE) return eval_coated_diffuse(C, m, wo, wi, ns, lam);
    if (TYPE == HK_MAT_DIFFUSE_TRANSMISSION) return eval_diffuse_transmission(C, m, wo, wi, ns, lam);
    if (TYPE == HK_MAT_COATED_CONDUCTOR) return eval_coated_conductor(C, m, wo, wi, ns, lam);
    if (TYPE == HK_MAT_COATED_DIFFUSE_TRANSMISSION) return eval_coated_difftrans(C, m, wo, wi, ns, lam);
    return eval_none();   // Mirror / Glass / ThinDielectric are delta-only
}
