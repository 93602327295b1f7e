// ok_bsdf_coated_conductor.h — CPU restatement of the reference's CoatedConductorMaterial (TEST INFRASTRUCTURE ONLY).
// Follows src/materials/spectral-eval.jl:2877-3237 (sample_bsdf_spectral) and :3243-3418 (evaluate_bsdf_spectral);
// parameters as in src/materials/coated-conductor.jl:48-105.  Despite the "LayeredBxDF" doc string the reference code is a
// closed form (one interface event + one conductor event, no random walk): that is what is restated, operand order kept.
#pragma once
// included from ok_bsdf.h after ok_bsdf_layered.h (layer_transmittance, the Trowbridge-Reitz helpers)

namespace ok {

struct CCParams {
    float ieta, iax, iay, cax, cay, thickness;
    Spec ce, ck, albedo;
    bool has_medium;
};
inline Spec spec_sqrt(const Spec& a) { return Spec(std::sqrt(a.v[0]), std::sqrt(a.v[1]), std::sqrt(a.v[2]), std::sqrt(a.v[3])); }
inline Spec spec_clamp_zero(const Spec& a) { return Spec(std::max(a.v[0], 0.0f), std::max(a.v[1], 0.0f), std::max(a.v[2], 0.0f), std::max(a.v[3], 0.0f)); }

// the parameter block both functions start with (:2891-2951 / :3263-3303)
inline CCParams coated_conductor_params(const MatCtx& C, const HkMaterial& m, const Wavelengths& l, bool regularize) {
    CCParams P;
    P.ieta = m.f[3];
    if (P.ieta == 0.0f) P.ieta = 1.0f;
    const bool remap = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) != 0;
    P.iax = remap ? roughness_to_alpha(m.f[0]) : m.f[0];
    P.iay = remap ? roughness_to_alpha(m.f[1]) : m.f[1];
    P.cax = remap ? roughness_to_alpha(m.f[5]) : m.f[5];
    P.cay = remap ? roughness_to_alpha(m.f[6]) : m.f[6];
    if (regularize) {
        P.iax = regularize_alpha(P.iax); P.iay = regularize_alpha(P.iay);
        P.cax = regularize_alpha(P.cax); P.cay = regularize_alpha(P.cay);
    }
    if (m.flags & HK_MATFLAG_USE_ETA_K) {
        P.ce = eval_ior_spectral(C, m, 0, l);
        P.ck = eval_ior_spectral(C, m, 1, l);
    } else {   // reflectance mode: eta = 1, k = 2 sqrt(r) / sqrt(1 - r)   (:2926-2939)
        float refl[3];
        for (int i = 0; i < 3; i++) refl[i] = clampf(m.rgb0[i], 0.0f, 0.9999f);
        Spec r = uplift_rgb(*C.T, refl, l);
        P.ce = Spec(1.0f);
        P.ck = 2.0f * spec_sqrt(r) / spec_sqrt(spec_clamp_zero(Spec(1.0f) - r) + Spec(1.0e-6f));
    }
    P.ce = P.ce / P.ieta;     // conductor eta / k relative to the coating (:2942-2943)
    P.ck = P.ck / P.ieta;
    P.thickness = std::max(m.f[2], 1.1920929e-7f);
    P.albedo = uplift_rgb(*C.T, m.rgb2, l);
    P.has_medium = !(m.rgb2[0] == 0.0f && m.rgb2[1] == 0.0f && m.rgb2[2] == 0.0f);
    return P;
}
inline Spec cc_layer_tr(const CCParams& P, float tr_a, float tr_b) { return P.has_medium ? (tr_a * tr_b) * P.albedo : Spec(1.0f); }

// spectral-eval.jl:2877-3237
inline BSDFSample sample_coated_conductor(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2 sample_u, float rng, bool regularize) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    CCParams P = coated_conductor_params(C, m, l, regularize);
    const float ieta = P.ieta;
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), wo_dot_n);
    const bool flip = wo_l.z < 0.0f;
    if (flip) wo_l = -wo_l;
    const float cos_o = std::fabs(wo_l.z);
    const bool i_smooth = tr_effectively_smooth(P.iax, P.iay), c_smooth = tr_effectively_smooth(P.cax, P.cay);
    auto to_world = [&](V3 wl) { if (flip) wl = -wl; return normalize(tg * wl.x + bt * wl.y + n * wl.z); };

    if (i_smooth) {
        const float F_i = fresnel_dielectric(cos_o, ieta);
        if (rng < F_i) return BSDFSample(to_world(V3(-wo_l.x, -wo_l.y, wo_l.z)), Spec(1.0f), 1.0f, true, 1.0f);
        const float sin2_t = std::max(0.0f, 1.0f - cos_o * cos_o) / (ieta * ieta);
        if (sin2_t >= 1.0f) return BSDFSample();
        const float cos_t_in = std::sqrt(1.0f - sin2_t);
        if (c_smooth) {
            V3 wi_base = normalize(V3(-wo_l.x / ieta, -wo_l.y / ieta, cos_t_in));
            Spec F_c = fr_complex_spectral(cos_t_in, P.ce, P.ck);
            const float sin2_out = std::max(0.0f, 1.0f - wi_base.z * wi_base.z) * (ieta * ieta);
            if (sin2_out >= 1.0f) return BSDFSample();
            const float cos_out = std::sqrt(1.0f - sin2_out);
            const float T_in = 1.0f - F_i, T_out = 1.0f - fresnel_dielectric(cos_out, ieta);
            const float tr = P.has_medium ? layer_transmittance(P.thickness, V3(0, 0, cos_t_in)) : 1.0f;
            Spec f = F_c * T_in * T_out * cc_layer_tr(P, tr, tr) / cos_o;
            return BSDFSample(to_world(V3(-wo_l.x, -wo_l.y, wo_l.z)), f, 1.0f - F_i, true, 1.0f);
        }
        V3 wo_c = normalize(V3(wo_l.x / ieta, wo_l.y / ieta, cos_t_in));
        const float cax = std::max(P.cax, 1.0e-4f), cay = std::max(P.cay, 1.0e-4f);
        V3 wm = tr_sample_wm(wo_c, sample_u, cax, cay);
        const float cos_om = dot(wo_c, wm);
        if (cos_om < 0.0f) return BSDFSample();
        V3 wi_c = -wo_c + 2.0f * cos_om * wm;
        if (wi_c.z < 0.0f) return BSDFSample();
        Spec F_c = fr_complex_spectral(std::fabs(cos_om), P.ce, P.ck);
        const float D = tr_d(wm, cax, cay), G = tr_g(wo_c, wi_c, cax, cay);
        Spec f_c = D * F_c * G / (4.0f * std::fabs(wo_c.z) * std::fabs(wi_c.z));
        const float sin2_out = (wi_c.x * wi_c.x + wi_c.y * wi_c.y) * (ieta * ieta);
        if (sin2_out >= 1.0f) return BSDFSample();
        const float cos_out = std::sqrt(1.0f - sin2_out);
        const float T_in = 1.0f - F_i, T_out = 1.0f - fresnel_dielectric(cos_out, ieta);
        Spec ltr = P.has_medium ? cc_layer_tr(P, layer_transmittance(P.thickness, V3(0, 0, cos_t_in)), layer_transmittance(P.thickness, V3(0, 0, wi_c.z))) : Spec(1.0f);
        V3 wi_l = normalize(V3(wi_c.x * ieta, wi_c.y * ieta, cos_out));
        Spec f = f_c * T_in * T_out * ltr;
        const float pdf_c = tr_pdf(wo_c, wm, cax, cay) / (4.0f * std::fabs(cos_om));
        return BSDFSample(to_world(wi_l), f, (1.0f - F_i) * pdf_c, false, 1.0f);
    }

    // rough coating
    const float iax = std::max(P.iax, 1.0e-4f), iay = std::max(P.iay, 1.0e-4f);
    V3 wm = tr_sample_wm(wo_l, sample_u, iax, iay);
    const float cos_om = dot(wo_l, wm);
    if (cos_om < 0.0f) return BSDFSample();
    const float F_i = fresnel_dielectric(cos_om, ieta);
    if (rng < F_i) {
        V3 wi_l = -wo_l + 2.0f * cos_om * wm;
        if (wi_l.z * wo_l.z < 0.0f) return BSDFSample();
        V3 wi = to_world(wi_l);
        if (flip) wi_l = -wi_l;
        const float D = tr_d(wm, iax, iay), G = tr_g(wo_l, wi_l, iax, iay);
        const float cos_i = std::fabs(wi_l.z);
        const float pdf = F_i * tr_pdf(wo_l, wm, iax, iay) / (4.0f * std::fabs(cos_om));
        const float f = D * G / (4.0f * cos_i * cos_o);
        return BSDFSample(wi, Spec(f), pdf, false, 1.0f);
    }
    const float T_in = 1.0f - F_i;
    if (c_smooth) {
        V3 lcw(-wo_l.x, -wo_l.y, wo_l.z);
        const float cos_b = std::fabs(lcw.z);
        Spec F_c = fr_complex_spectral(cos_b, P.ce, P.ck);
        const float T_out = 1.0f - fresnel_dielectric(cos_b, ieta);
        const float tr = P.has_medium ? layer_transmittance(P.thickness, lcw) : 1.0f;
        Spec f = F_c * T_in * T_out * cc_layer_tr(P, tr, tr) / cos_o;
        const float pdf = (1.0f - F_i) * tr_pdf(wo_l, wm, iax, iay) / (4.0f * std::fabs(cos_om));
        return BSDFSample(to_world(lcw), f, pdf, false, 1.0f);
    }
    const float cax = std::max(P.cax, 1.0e-4f), cay = std::max(P.cay, 1.0e-4f);
    V3 wm_c = tr_sample_wm(wo_l, sample_u, cax, cay);
    const float cos_omc = dot(wo_l, wm_c);
    if (cos_omc < 0.0f) return BSDFSample();
    V3 wi_l = -wo_l + 2.0f * cos_omc * wm_c;
    if (wi_l.z * wo_l.z < 0.0f) return BSDFSample();
    Spec F_c = fr_complex_spectral(std::fabs(cos_omc), P.ce, P.ck);
    const float D = tr_d(wm_c, cax, cay), G = tr_g(wo_l, wi_l, cax, cay);
    const float cos_i = std::fabs(wi_l.z);
    Spec f_c = D * F_c * G / (4.0f * cos_i * cos_o);
    const float T_out = 1.0f - fresnel_dielectric(cos_i, ieta);
    Spec ltr = P.has_medium ? cc_layer_tr(P, layer_transmittance(P.thickness, V3(0, 0, cos_o)), layer_transmittance(P.thickness, wi_l)) : Spec(1.0f);
    Spec f = f_c * T_in * T_out * ltr;
    const float pdf = (1.0f - F_i) * tr_pdf(wo_l, wm_c, cax, cay) / (4.0f * std::fabs(cos_omc));
    return BSDFSample(to_world(wi_l), f, pdf, false, 1.0f);
}

// spectral-eval.jl:3243-3418
inline BSDFEval eval_coated_conductor(const MatCtx& C, const HkMaterial& m, V3 wo, V3 wi, V3 n, const Wavelengths& l) {
    const float cos_i = dot(wi, n), cos_o = dot(wo, n);
    if (cos_i * cos_o < 0.0f) return BSDFEval();
    if (std::fabs(cos_i) < 1.0e-6f || std::fabs(cos_o) < 1.0e-6f) return BSDFEval();
    CCParams P = coated_conductor_params(C, m, l, false);
    const float ieta = P.ieta;
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), cos_o), wi_l(dot(wi, tg), dot(wi, bt), cos_i);
    if (wo_l.z < 0.0f) { wo_l = -wo_l; wi_l = -wi_l; }
    const bool i_smooth = tr_effectively_smooth(P.iax, P.iay), c_smooth = tr_effectively_smooth(P.cax, P.cay);
    if (i_smooth && c_smooth) return BSDFEval();
    V3 wh = normalize(wo_l + wi_l);
    if (wh.z < 0.0f) wh = -wh;
    const float cos_oh = dot(wo_l, wh);
    const float F_wh = fresnel_dielectric(std::fabs(cos_oh), ieta);
    const float T_o = 1.0f - fresnel_dielectric(std::fabs(wo_l.z), ieta), T_i = 1.0f - fresnel_dielectric(std::fabs(wi_l.z), ieta);
    const float tr = P.has_medium ? layer_transmittance(P.thickness, wi_l) : 1.0f;
    Spec ltr = cc_layer_tr(P, tr, tr);
    const float denom = 4.0f * std::fabs(wi_l.z) * std::fabs(wo_l.z);
    if (i_smooth) {
        const float cax = std::max(P.cax, 1.0e-4f), cay = std::max(P.cay, 1.0e-4f);
        const float D = tr_d(wh, cax, cay), G = tr_g(wo_l, wi_l, cax, cay);
        Spec F_c = fr_complex_spectral(std::fabs(cos_oh), P.ce, P.ck);
        Spec f_c = D * F_c * G / denom;
        Spec f = f_c * T_o * T_i * ltr;
        const float pdf = T_o * tr_pdf(wo_l, wh, cax, cay) / (4.0f * std::fabs(cos_oh));
        return BSDFEval(f, pdf);
    }
    const float iax = std::max(P.iax, 1.0e-4f), iay = std::max(P.iay, 1.0e-4f);
    const float D_i = tr_d(wh, iax, iay), G_i = tr_g(wo_l, wi_l, iax, iay);
    const float f_interface = D_i * F_wh * G_i / denom;
    Spec f_c; float pdf_c;
    if (c_smooth) {
        f_c = fr_complex_spectral(std::fabs(wo_l.z), P.ce, P.ck) / std::fabs(wo_l.z);
        pdf_c = 1.0f;
    } else {
        const float cax = std::max(P.cax, 1.0e-4f), cay = std::max(P.cay, 1.0e-4f);
        const float D_c = tr_d(wh, cax, cay), G_c = tr_g(wo_l, wi_l, cax, cay);
        Spec F_c = fr_complex_spectral(std::fabs(cos_oh), P.ce, P.ck);
        f_c = D_c * F_c * G_c / denom;
        pdf_c = tr_pdf(wo_l, wh, cax, cay) / (4.0f * std::fabs(cos_oh));
    }
    Spec f = Spec(f_interface) + f_c * T_o * T_i * ltr;
    const float F_io = fresnel_dielectric(std::fabs(wo_l.z), ieta);
    const float pdf_interface = F_io * tr_pdf(wo_l, wh, iax, iay) / (4.0f * std::fabs(cos_oh));
    return BSDFEval(f, pdf_interface + T_o * pdf_c);
}

}  // namespace ok
