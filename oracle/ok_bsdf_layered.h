// ok_bsdf_layered.h — ORACLE (test infrastructure, NOT product code).
// CoatedDiffuse (LayeredBxDF random walk), ThinDielectric, DiffuseTransmission.
// Restates src/materials/spectral-eval.jl:823-2218 literally, including its quirks
// (e.g. the argument order of lerp() at :1936 and `phase_p / phase_p` at :1712).
#pragma once
// (included from ok_bsdf.h inside no namespace)
namespace ok {

// spectral-eval.jl:823-825
inline float sample_exponential(float u, float a) { return -dm_logf(1.0f - u) / a; }
// spectral-eval.jl:837-840
inline float layer_transmittance(float thickness, V3 w) {
    if (std::fabs(thickness) <= 1.1920929e-7f) return 1.0f;
    return dm_expf(-std::fabs(thickness / w.z));
}
// spectral-eval.jl:879-883
inline float hg_phase_pdf(float g, float cos_t) {
    float g2 = g * g;
    float denom = 1.0f + g2 - 2.0f * g * cos_t;
    return (1.0f - g2) / (4.0f * PI_F * denom * std::sqrt(std::max(1.0e-10f, denom)));
}
// spectral-eval.jl:847-872
inline V3 sample_hg_phase_spectral(float g, V3 wo, V2 u, float& p_out) {
    float cos_t;
    if (std::fabs(g) < 1.0e-3f) cos_t = 1.0f - 2.0f * u.x;
    else {
        float g2 = g * g;
        float sq = (1.0f - g2) / (1.0f - g + 2.0f * g * u.x);
        cos_t = clampf((1.0f + g2 - sq * sq) / (2.0f * g), -1.0f, 1.0f);
    }
    float sin_t = std::sqrt(std::max(0.0f, 1.0f - cos_t * cos_t));
    float phi = 2.0f * PI_F * u.y;
    V3 t1, t2; coordinate_system(-wo, t1, t2);
    V3 wi = sin_t * dm_cosf(phi) * t1 + sin_t * dm_sinf(phi) * t2 + cos_t * (-wo);
    wi = normalize(wi);
    float g2 = g * g;
    float denom = 1.0f + g2 - 2.0f * g * cos_t;
    p_out = (1.0f - g2) / (4.0f * PI_F * denom * std::sqrt(std::max(1.0e-10f, denom)));
    return wi;
}

static constexpr uint8_t BXDF_REFLECTION = 1, BXDF_TRANSMISSION = 2, BXDF_ALL = 3;

struct LSample {   // LayeredBSDFSample, spectral-eval.jl:953-963
    Spec f; V3 wi; float pdf; bool is_reflection, is_specular; float eta; bool valid;
    LSample() : f(), wi(0, 0, 0), pdf(0.0f), is_reflection(false), is_specular(false), eta(1.0f), valid(false) {}
    LSample(Spec ff, V3 w, float p, bool r, bool s, float e, bool v) : f(ff), wi(w), pdf(p), is_reflection(r), is_specular(s), eta(e), valid(v) {}
};

// spectral-eval.jl:1072-1093
inline bool refract_pbrt(V3 wo, float eta, V3& wi, float& etap) {
    float ci = wo.z;
    etap = ci > 0.0f ? eta : (1.0f / eta);
    float s2i = std::max(0.0f, 1.0f - ci * ci);
    float s2t = s2i / (etap * etap);
    if (s2t >= 1.0f) { wi = V3(0, 0, 0); etap = 1.0f; return false; }
    float ct = std::sqrt(1.0f - s2t);
    float cts = ci > 0.0f ? -ct : ct;
    wi = normalize(V3(-wo.x / etap, -wo.y / etap, cts));
    return true;
}
// spectral-eval.jl:1100-1120
inline bool refract_microfacet(V3 wo, V3 wm, float eta, V3& wi, float& etap) {
    float ci = dot(wo, wm);
    etap = ci > 0.0f ? eta : (1.0f / eta);
    float s2i = std::max(0.0f, 1.0f - ci * ci);
    float s2t = s2i / (etap * etap);
    if (s2t >= 1.0f) { wi = V3(0, 0, 0); etap = 1.0f; return false; }
    float ct = std::sqrt(1.0f - s2t);
    float cts = ci > 0.0f ? -ct : ct;
    wi = normalize(-wo / etap + (ci / etap + cts) * wm);
    return true;
}

// spectral-eval.jl:973-1063
inline LSample sample_dielectric_interface(V3 wo, float uc, V2 u, float ax, float ay, float eta, uint8_t flags) {
    bool smooth = tr_effectively_smooth(ax, ay);
    if (smooth || eta == 1.0f) {
        float R = fresnel_dielectric(wo.z, eta), T = 1.0f - R;
        float pr = (flags & BXDF_REFLECTION) ? R : 0.0f;
        float pt = (flags & BXDF_TRANSMISSION) ? T : 0.0f;
        if (pr == 0.0f && pt == 0.0f) return LSample();
        if (uc < pr / (pr + pt)) {
            V3 wi(-wo.x, -wo.y, wo.z);
            return LSample(Spec(R / std::fabs(wi.z)), wi, pr / (pr + pt), true, true, 1.0f, true);
        }
        V3 wi; float etap;
        if (!refract_pbrt(wo, eta, wi, etap)) return LSample();
        return LSample(Spec(T / std::fabs(wi.z)), wi, pt / (pr + pt), false, true, etap, true);
    }
    V3 wm = tr_sample_wm(wo, u, ax, ay);
    float com = dot(wo, wm);
    float R = fresnel_dielectric(com, eta), T = 1.0f - R;
    float pr = (flags & BXDF_REFLECTION) ? R : 0.0f;
    float pt = (flags & BXDF_TRANSMISSION) ? T : 0.0f;
    if (pr == 0.0f && pt == 0.0f) return LSample();
    if (uc < pr / (pr + pt)) {
        V3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return LSample();
        float pdf_m = tr_pdf(wo, wm, ax, ay);
        float pdf = pdf_m / (4.0f * std::fabs(com)) * pr / (pr + pt);
        float D = tr_d(wm, ax, ay), G = tr_g(wo, wi, ax, ay);
        float f = D * G * R / (4.0f * wo.z * wi.z);
        return LSample(Spec(f), wi, pdf, true, false, 1.0f, true);
    }
    V3 wi; float etap;
    bool valid = refract_microfacet(wo, wm, eta, wi, etap);
    if (!valid || same_hemisphere(wo, wi) || wi.z == 0.0f) return LSample();
    float dd = dot(wi, wm) + dot(wo, wm) / etap;
    float denom = dd * dd;
    float dwm_dwi = std::fabs(dot(wi, wm)) / denom;
    float pdf_m = tr_pdf(wo, wm, ax, ay);
    float pdf = pdf_m * dwm_dwi * pt / (pr + pt);
    float D = tr_d(wm, ax, ay), G = tr_g(wo, wi, ax, ay);
    float f = T * D * G * std::fabs(dot(wi, wm) * dot(wo, wm) / (wi.z * wo.z * denom));
    return LSample(Spec(f), wi, pdf, false, false, etap, true);
}
// spectral-eval.jl:1144-1171
inline LSample sample_diffuse_interface(V3 wo, V2 u, const Spec& refl, uint8_t flags) {
    if ((flags & BXDF_REFLECTION) == 0) return LSample();
    V3 wi = cosine_sample_hemisphere(u);
    if (wo.z < 0.0f) wi = V3(wi.x, wi.y, -wi.z);
    float ci = std::fabs(wi.z);
    if (ci < 1.0e-6f) return LSample();
    return LSample(refl * (1.0f / PI_F), wi, ci / PI_F, true, false, 1.0f, true);
}
// spectral-eval.jl:1178-1199
inline Spec eval_diffuse_interface(V3 wo, V3 wi, const Spec& refl, float* pdf = nullptr) {
    if (!same_hemisphere(wo, wi)) { if (pdf) *pdf = 0.0f; return Spec(); }
    if (pdf) *pdf = std::fabs(wi.z) / PI_F;
    return refl * (1.0f / PI_F);
}
inline float pdf_diffuse_interface(V3 wo, V3 wi) { return same_hemisphere(wo, wi) ? std::fabs(wi.z) / PI_F : 0.0f; }
// spectral-eval.jl:1206-1215
inline float power_heuristic(int nf, float fp, int ng, float gp) {
    float f = (float)nf * fp, g = (float)ng * gp;
    float f2 = f * f, g2 = g * g;
    if (f2 + g2 == 0.0f) return 0.0f;
    return f2 / (f2 + g2);
}
// spectral-eval.jl:1426-1486  (returns f; the pdf it also computes is unused by every caller)
inline Spec eval_dielectric_interface(V3 wo, V3 wi, float ax, float ay, float eta) {
    if (tr_effectively_smooth(ax, ay) || eta == 1.0f) return Spec();
    if (same_hemisphere(wo, wi)) {
        V3 wh = normalize(wo + wi);
        if (wh.z < 0.0f) wh = -wh;
        float coh = dot(wo, wh);
        float R = fresnel_dielectric(coh, eta);
        float D = tr_d(wh, ax, ay), G = tr_g(wo, wi, ax, ay);
        return Spec(D * G * R / (4.0f * wo.z * wi.z));
    }
    float etap = wo.z > 0.0f ? eta : (1.0f / eta);
    V3 wh = normalize(wo + wi * etap);
    if (wh.z < 0.0f) wh = -wh;
    float coh = dot(wo, wh), cih = dot(wi, wh);
    if (coh * cih > 0.0f) return Spec();
    float R = fresnel_dielectric(coh, eta), T = 1.0f - R;
    float dd = cih + coh / etap;
    float denom = dd * dd;
    float D = tr_d(wh, ax, ay), G = tr_g(wo, wi, ax, ay);
    return Spec(T * D * G * std::fabs(cih * coh / (wo.z * wi.z * denom)));
}
// spectral-eval.jl:1493-1554
inline float pdf_dielectric_interface(V3 wo, V3 wi, float ax, float ay, float eta, uint8_t flags = BXDF_ALL) {
    if (tr_effectively_smooth(ax, ay) || eta == 1.0f) return 0.0f;
    if (same_hemisphere(wo, wi)) {
        if ((flags & BXDF_REFLECTION) == 0) return 0.0f;
        V3 wh = normalize(wo + wi);
        if (wh.z < 0.0f) wh = -wh;
        float coh = std::fabs(dot(wo, wh));
        float R = fresnel_dielectric(coh, eta), T = 1.0f - R;
        float pr = (flags & BXDF_REFLECTION) ? R : 0.0f;
        float pt = (flags & BXDF_TRANSMISSION) ? T : 0.0f;
        float pdf = tr_pdf(wo, wh, ax, ay) / (4.0f * coh);
        return pdf * pr / (pr + pt);
    }
    if ((flags & BXDF_TRANSMISSION) == 0) return 0.0f;
    float etap = wo.z > 0.0f ? eta : (1.0f / eta);
    V3 wh = normalize(wo + wi * etap);
    if (wh.z < 0.0f) wh = -wh;
    float coh = dot(wo, wh), cih = dot(wi, wh);
    if (coh * cih > 0.0f) return 0.0f;
    float R = fresnel_dielectric(std::fabs(coh), eta), T = 1.0f - R;
    float pr = (flags & BXDF_REFLECTION) ? R : 0.0f;
    float pt = (flags & BXDF_TRANSMISSION) ? T : 0.0f;
    float dd = cih + coh / etap;
    float denom = dd * dd;
    float dwm_dwi = std::fabs(cih) / denom;
    float pdf = tr_pdf(wo, wh, ax, ay) * dwm_dwi;
    return pdf * pt / (pr + pt);
}

struct CoatedParams {
    float refl_rgb[3], albedo_rgb[3];
    float eta, thickness, g, ax, ay;
    int max_depth, n_samples;
    bool has_medium;
};
inline CoatedParams coated_params(const HkMaterial& m, bool regularize) {
    CoatedParams p;
    for (int i = 0; i < 3; i++) { p.refl_rgb[i] = m.rgb0[i]; p.albedo_rgb[i] = m.rgb1[i]; }
    p.eta = m.f[3];
    p.thickness = std::max(m.f[2], 1.1920929e-7f);
    p.g = clampf(m.f[4], -0.99f, 0.99f);
    bool remap = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) != 0;
    p.ax = remap ? roughness_to_alpha(m.f[0]) : m.f[0];
    p.ay = remap ? roughness_to_alpha(m.f[1]) : m.f[1];
    if (regularize) { p.ax = regularize_alpha(p.ax); p.ay = regularize_alpha(p.ay); }
    p.max_depth = m.ival[0];
    p.n_samples = m.ival[1];
    p.has_medium = !(m.rgb1[0] == 0.0f && m.rgb1[1] == 0.0f && m.rgb1[2] == 0.0f);
    return p;
}

// spectral-eval.jl:1232-1418
inline BSDFSample sample_coated_diffuse(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2 sample_u, float rng_in, bool regularize) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    CoatedParams P = coated_params(m, regularize);
    Spec refl = uplift_rgb(*C.T, P.refl_rgb, l);
    Spec albedo = uplift_rgb(*C.T, P.albedo_rgb, l);
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), wo_dot_n);
    bool flip = wo_l.z < 0.0f;
    if (flip) wo_l = -wo_l;
    const float thickness = P.thickness;
    LSample bs = sample_dielectric_interface(wo_l, rng_in, sample_u, P.ax, P.ay, P.eta, BXDF_ALL);
    if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) return BSDFSample();
    if (bs.is_reflection) {
        V3 wl = bs.wi; if (flip) wl = -wl;
        V3 wi = normalize(tg * wl.x + bt * wl.y + n * wl.z);
        return BSDFSample(wi, bs.f, bs.pdf, bs.is_specular, 1.0f);
    }
    V3 w = bs.wi;
    bool specular_path = bs.is_specular;
    Spec f = bs.f * std::fabs(w.z);
    float pdf = bs.pdf;
    float z = thickness;
    PCG32 rng = pcg32_init(pbrt_hash((uint64_t)0, wo_l), pbrt_hash(rng_in, sample_u));
    for (int depth = 0; depth < P.max_depth; depth++) {
        float rr_beta = max_component(f) / pdf;
        if (depth > 3 && rr_beta < 0.25f) {
            float q = std::max(0.0f, 1.0f - rr_beta);
            float rv = pcg32_f32(rng);
            if (rv < q) return BSDFSample();
            pdf *= 1.0f - q;
        }
        if (w.z == 0.0f) return BSDFSample();
        if (P.has_medium) {
            float eu = pcg32_f32(rng);
            float dz = sample_exponential(eu, 1.0f / std::fabs(w.z));
            float zp = w.z > 0.0f ? (z + dz) : (z - dz);
            if (zp == z) return BSDFSample();
            if (0.0f < zp && zp < thickness) {
                float p1 = pcg32_f32(rng), p2 = pcg32_f32(rng);
                float phase_p;
                V3 wip = sample_hg_phase_spectral(P.g, -w, V2(p1, p2), phase_p);
                if (phase_p == 0.0f || wip.z == 0.0f) return BSDFSample();
                f = f * albedo * phase_p;
                pdf *= phase_p;
                specular_path = false;
                w = wip;
                z = zp;
                continue;
            }
            z = clampf(zp, 0.0f, thickness);
        } else {
            z = (z == thickness) ? 0.0f : thickness;
            f = f * layer_transmittance(thickness, w);
        }
        bool at_bottom = z == 0.0f;
        float uc = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
        LSample bi = at_bottom ? sample_diffuse_interface(-w, V2(u1, u2), refl, BXDF_ALL)
                               : sample_dielectric_interface(-w, uc, V2(u1, u2), P.ax, P.ay, P.eta, BXDF_ALL);
        if (!bi.valid || bi.pdf == 0.0f || bi.wi.z == 0.0f) return BSDFSample();
        f = f * bi.f;
        pdf *= bi.pdf;
        specular_path = specular_path && bi.is_specular;
        w = bi.wi;
        if (!bi.is_reflection) {
            V3 wl = w; if (flip) wl = -wl;
            V3 wi = normalize(tg * wl.x + bt * wl.y + n * wl.z);
            return BSDFSample(wi, f, pdf, specular_path, bi.eta);
        }
        f = f * std::fabs(bi.wi.z);
    }
    return BSDFSample();
}

// spectral-eval.jl:1848-1937
inline float pdf_layered_bsdf(V3 wo, V3 wi, float ax, float ay, float eta, int n_samples, int /*max_depth*/, const Spec& refl, bool, float, float) {
    PCG32 rng = pcg32_init(pbrt_hash((uint64_t)0, wi), pbrt_hash(wo));
    bool same_hemi = same_hemisphere(wo, wi);
    bool smooth = tr_effectively_smooth(ax, ay);
    float pdf_sum = 0.0f;
    if (same_hemi) {
        if (smooth) pdf_sum += (float)n_samples * 0.0f;
        else pdf_sum += (float)n_samples * pdf_dielectric_interface(wo, wi, ax, ay, eta, BXDF_REFLECTION);
    }
    for (int s = 0; s < n_samples; s++) {
        if (same_hemi) {
            float uc1 = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
            LSample wos = sample_dielectric_interface(wo, uc1, V2(u1, u2), ax, ay, eta, BXDF_TRANSMISSION);
            float uc2 = pcg32_f32(rng), u3 = pcg32_f32(rng), u4 = pcg32_f32(rng);
            LSample wis = sample_dielectric_interface(wi, uc2, V2(u3, u4), ax, ay, eta, BXDF_TRANSMISSION);
            if (wos.valid && wos.pdf > 0.0f && wis.valid && wis.pdf > 0.0f) {
                if (smooth) pdf_sum += pdf_diffuse_interface(-wos.wi, -wis.wi);
                else {
                    float u5 = pcg32_f32(rng), u6 = pcg32_f32(rng);
                    LSample rs = sample_diffuse_interface(-wos.wi, V2(u5, u6), refl, BXDF_ALL);
                    if (rs.valid && rs.pdf > 0.0f) {
                        float r_pdf = pdf_diffuse_interface(-wos.wi, -wis.wi);
                        float wt = power_heuristic(1, wis.pdf, 1, r_pdf);
                        pdf_sum += wt * r_pdf;
                        float t_pdf = pdf_dielectric_interface(-rs.wi, wi, ax, ay, eta);
                        float wt2 = power_heuristic(1, rs.pdf, 1, t_pdf);
                        pdf_sum += wt2 * t_pdf;
                    }
                }
            }
        } else {
            float uc1 = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
            LSample wos = sample_dielectric_interface(wo, uc1, V2(u1, u2), ax, ay, eta, BXDF_TRANSMISSION);
            if (!wos.valid || wos.pdf == 0.0f || wos.is_reflection) continue;
            float u3 = pcg32_f32(rng), u4 = pcg32_f32(rng);
            LSample wis = sample_diffuse_interface(wi, V2(u3, u4), refl, BXDF_TRANSMISSION);
            if (!wis.valid || wis.pdf == 0.0f || wis.is_reflection) continue;
            if (smooth) pdf_sum += pdf_diffuse_interface(-wos.wi, wi);
            else pdf_sum += (pdf_dielectric_interface(wo, -wis.wi, ax, ay, eta) + pdf_diffuse_interface(-wos.wi, wi)) / 2.0f;
        }
    }
    // literal: lerp(0.9f0, 1/(4π), pdf_sum/n) with lerp(v1,v2,t) = (1-t)*v1 + t*v2
    return lerpf(0.9f, 1.0f / (4.0f * PI_F), pdf_sum / (float)n_samples);
}

// spectral-eval.jl:1564-1840
inline BSDFEval eval_coated_diffuse(const MatCtx& C, const HkMaterial& m, V3 wo, V3 wi, V3 n, const Wavelengths& l) {
    CoatedParams P = coated_params(m, false);
    Spec refl = uplift_rgb(*C.T, P.refl_rgb, l);
    Spec albedo = uplift_rgb(*C.T, P.albedo_rgb, l);
    const float thickness = P.thickness, ax = P.ax, ay = P.ay, eta = P.eta, g = P.g;
    V3 tg, bt; coordinate_system(n, tg, bt);
    float co = dot(wo, n), ci = dot(wi, n);
    V3 wo_l(dot(wo, tg), dot(wo, bt), co), wi_l(dot(wi, tg), dot(wi, bt), ci);
    if (wo_l.z < 0.0f) { wo_l = -wo_l; wi_l = -wi_l; }
    if (std::fabs(wo_l.z) < 1.0e-6f || std::fabs(wi_l.z) < 1.0e-6f) return BSDFEval();
    bool same_hemi = same_hemisphere(wo_l, wi_l);
    bool exit_at_bottom = same_hemi ^ true;
    float exit_z = exit_at_bottom ? 0.0f : thickness;
    Spec fr;
    if (same_hemi) fr = fr + eval_dielectric_interface(wo_l, wi_l, ax, ay, eta) * (float)P.n_samples;
    PCG32 rng = pcg32_init(pbrt_hash((uint64_t)0, wo_l), pbrt_hash(wi_l));
    bool smooth = tr_effectively_smooth(ax, ay);
    for (int s = 0; s < P.n_samples; s++) {
        float uc = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
        LSample wos = sample_dielectric_interface(wo_l, uc, V2(u1, u2), ax, ay, eta, BXDF_TRANSMISSION);
        if (!wos.valid || wos.pdf == 0.0f || wos.wi.z == 0.0f) continue;
        uc = pcg32_f32(rng); u1 = pcg32_f32(rng); u2 = pcg32_f32(rng);
        LSample wis = exit_at_bottom ? sample_diffuse_interface(wi_l, V2(u1, u2), refl, BXDF_TRANSMISSION)
                                     : sample_dielectric_interface(wi_l, uc, V2(u1, u2), ax, ay, eta, BXDF_TRANSMISSION);
        if (!wis.valid || wis.pdf == 0.0f || wis.wi.z == 0.0f) continue;
        Spec beta = wos.f * std::fabs(wos.wi.z) / wos.pdf;
        float z = thickness;
        V3 w = wos.wi;
        for (int depth = 0; depth < P.max_depth; depth++) {
            if (depth > 3 && max_component(beta) < 0.25f) {
                float q = std::max(0.0f, 1.0f - max_component(beta));
                float rv = pcg32_f32(rng);
                if (rv < q) break;
                beta = beta / (1.0f - q);
            }
            if (P.has_medium) {
                float eu = pcg32_f32(rng);
                float dz = sample_exponential(eu, 1.0f / std::fabs(w.z));
                float zp = w.z > 0.0f ? (z + dz) : (z - dz);
                if (zp == z) continue;
                if (0.0f < zp && zp < thickness) {
                    float wt;
                    if (exit_at_bottom) {
                        wt = power_heuristic(1, wis.pdf, 1, hg_phase_pdf(g, dot(-w, -wis.wi)));
                    } else {
                        wt = !smooth ? power_heuristic(1, wis.pdf, 1, hg_phase_pdf(g, dot(-w, -wis.wi))) : 1.0f;
                    }
                    float phase_val = hg_phase_pdf(g, dot(-w, -wis.wi));
                    fr = fr + beta * albedo * phase_val * wt * layer_transmittance(zp - exit_z, wis.wi) * wis.f / wis.pdf;
                    float p1 = pcg32_f32(rng), p2 = pcg32_f32(rng);
                    float phase_p;
                    V3 wip = sample_hg_phase_spectral(g, -w, V2(p1, p2), phase_p);
                    if (phase_p == 0.0f || wip.z == 0.0f) break;
                    beta = beta * albedo * phase_p / phase_p;
                    w = wip;
                    z = zp;
                    if ((z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f)) {
                        Spec fe2; float exit_pdf;
                        if (exit_at_bottom) {
                            fe2 = eval_diffuse_interface(-w, wi_l, refl, &exit_pdf);
                        } else {
                            if (!smooth) {
                                fe2 = eval_dielectric_interface(-w, wi_l, ax, ay, eta);
                                exit_pdf = pdf_dielectric_interface(-w, wi_l, ax, ay, eta, BXDF_TRANSMISSION);
                            } else continue;
                        }
                        if (max_component(fe2) > 0.0f) {
                            float wt2 = power_heuristic(1, phase_p, 1, exit_pdf);
                            fr = fr + beta * layer_transmittance(zp - exit_z, wip) * fe2 * wt2;
                        }
                    }
                    continue;
                }
                z = clampf(zp, 0.0f, thickness);
            } else {
                z = (z == thickness) ? 0.0f : thickness;
                beta = beta * layer_transmittance(thickness, w);
            }
            bool at_exit = z == exit_z;
            if (at_exit) {
                float uc2 = pcg32_f32(rng), v1 = pcg32_f32(rng), v2 = pcg32_f32(rng);
                LSample bs = exit_at_bottom ? sample_diffuse_interface(-w, V2(v1, v2), refl, BXDF_REFLECTION)
                                            : sample_dielectric_interface(-w, uc2, V2(v1, v2), ax, ay, eta, BXDF_REFLECTION);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * std::fabs(bs.wi.z) / bs.pdf;
                w = bs.wi;
            } else {
                bool non_exit_specular = (z == thickness) ? smooth : false;
                if (!non_exit_specular) {
                    Spec f_nee = (z == thickness) ? eval_dielectric_interface(-w, -wis.wi, ax, ay, eta)
                                                  : eval_diffuse_interface(-w, -wis.wi, refl);
                    if (max_component(f_nee) > 0.0f) {
                        float wt = 1.0f;
                        if (!exit_at_bottom || !smooth) {
                            float nee_pdf = (z == thickness) ? pdf_dielectric_interface(-w, -wis.wi, ax, ay, eta)
                                                             : pdf_diffuse_interface(-w, -wis.wi);
                            wt = power_heuristic(1, wis.pdf, 1, nee_pdf);
                        }
                        fr = fr + beta * f_nee * std::fabs(wis.wi.z) * wt * layer_transmittance(thickness, wis.wi) * wis.f / wis.pdf;
                    }
                }
                float uc2 = pcg32_f32(rng), v1 = pcg32_f32(rng), v2 = pcg32_f32(rng);
                LSample bs = (z == thickness) ? sample_dielectric_interface(-w, uc2, V2(v1, v2), ax, ay, eta, BXDF_REFLECTION)
                                              : sample_diffuse_interface(-w, V2(v1, v2), refl, BXDF_REFLECTION);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * std::fabs(bs.wi.z) / bs.pdf;
                w = bs.wi;
                if (!smooth || exit_at_bottom) {
                    Spec fe3 = exit_at_bottom ? eval_diffuse_interface(-w, wi_l, refl)
                                              : eval_dielectric_interface(-w, wi_l, ax, ay, eta);
                    if (max_component(fe3) > 0.0f) {
                        float wt3 = 1.0f;
                        if (!non_exit_specular) {
                            float ep3 = exit_at_bottom ? pdf_diffuse_interface(-w, wi_l)
                                                       : pdf_dielectric_interface(-w, wi_l, ax, ay, eta, BXDF_TRANSMISSION);
                            wt3 = power_heuristic(1, bs.pdf, 1, ep3);
                        }
                        fr = fr + beta * layer_transmittance(thickness, bs.wi) * fe3 * wt3;
                    }
                }
            }
        }
    }
    fr = fr / (float)P.n_samples;
    float pdf = pdf_layered_bsdf(wo_l, wi_l, ax, ay, eta, P.n_samples, P.max_depth, refl, P.has_medium, g, thickness);
    return BSDFEval(fr, pdf);
}

// ---------------------------------------------------------------------------------------------
// ThinDielectric  spectral-eval.jl:1975-2037
// ---------------------------------------------------------------------------------------------
inline BSDFSample sample_thin_dielectric(const MatCtx&, const HkMaterial& m, V3 wo, V3 n, const Wavelengths&, V2, float rng, bool) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    float eta = m.f[0];
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), wo_dot_n);
    float co = std::fabs(wo_l.z);
    float R0 = fresnel_dielectric(co, eta), T0 = 1.0f - R0;
    float R = R0;
    if (R0 < 1.0f) R = R0 + T0 * T0 * R0 / (1.0f - R0 * R0);
    float T = 1.0f - R;
    float pr = R, pt = T;
    if (pr + pt < 1.0e-10f) return BSDFSample();
    float prob_reflect = pr / (pr + pt);
    if (rng < prob_reflect) {
        V3 wl(-wo_l.x, -wo_l.y, wo_l.z);
        V3 wi = normalize(tg * wl.x + bt * wl.y + n * wl.z);
        return BSDFSample(wi, Spec(R / std::fabs(wl.z)), prob_reflect, true, 1.0f);
    }
    return BSDFSample(-wo, Spec(T / co), 1.0f - prob_reflect, true, 1.0f);
}

// ---------------------------------------------------------------------------------------------
// DiffuseTransmission  spectral-eval.jl:2083-2218
// ---------------------------------------------------------------------------------------------
inline void difftrans_rgb(const HkMaterial& m, float* r, float* t) {
    float scale = m.f[0];
    for (int i = 0; i < 3; i++) { r[i] = clampf(m.rgb0[i] * scale, 0.0f, 1.0f); t[i] = clampf(m.rgb1[i] * scale, 0.0f, 1.0f); }
}
inline BSDFSample sample_diffuse_transmission(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2 u, float rng, bool) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    float r[3], t[3]; difftrans_rgb(m, r, t);
    Spec rs = uplift_rgb(*C.T, r, l), ts = uplift_rgb(*C.T, t, l);
    float pr = std::max(std::max(r[0], r[1]), r[2]), pt = std::max(std::max(t[0], t[1]), t[2]);
    if (pr + pt < 1.0e-10f) return BSDFSample();
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), wo_dot_n);
    float prob_reflect = pr / (pr + pt);
    V3 lw = cosine_sample_hemisphere(u);
    if (rng < prob_reflect) {
        if (wo_l.z < 0.0f) lw = V3(lw.x, lw.y, -lw.z);
        float c = std::fabs(lw.z);
        if (c < 1.0e-6f) return BSDFSample();
        V3 wi = normalize(tg * lw.x + bt * lw.y + n * lw.z);
        return BSDFSample(wi, rs * (1.0f / PI_F), prob_reflect * c / PI_F, false, 1.0f);
    }
    if (wo_l.z > 0.0f) lw = V3(lw.x, lw.y, -lw.z);
    float c = std::fabs(lw.z);
    if (c < 1.0e-6f) return BSDFSample();
    V3 wi = normalize(tg * lw.x + bt * lw.y + n * lw.z);
    return BSDFSample(wi, ts * (1.0f / PI_F), (1.0f - prob_reflect) * c / PI_F, false, 1.0f);
}
inline BSDFEval eval_diffuse_transmission(const MatCtx& C, const HkMaterial& m, V3 wo, V3 wi, V3 n, const Wavelengths& l) {
    float ci = dot(wi, n), co = dot(wo, n);
    float aci = std::fabs(ci);
    if (aci < 1.0e-6f) return BSDFEval();
    float r[3], t[3]; difftrans_rgb(m, r, t);
    Spec rs = uplift_rgb(*C.T, r, l), ts = uplift_rgb(*C.T, t, l);
    float pr = std::max(std::max(r[0], r[1]), r[2]), pt = std::max(std::max(t[0], t[1]), t[2]);
    if (pr + pt < 1.0e-10f) return BSDFEval();
    if (ci * co > 0.0f) return BSDFEval(rs * (1.0f / PI_F), pr / (pr + pt) * aci / PI_F);
    return BSDFEval(ts * (1.0f / PI_F), pt / (pr + pt) * aci / PI_F);
}

}  // namespace ok
