// hk_bsdf_coated_conductor.cuh — CoatedConductorMaterial: dielectric coating over a conductor.
// Reference: src/materials/spectral-eval.jl:2877-3237 (sample), :3243-3418 (eval); src/materials/coated-conductor.jl:48-105.
// The reference evaluates this material in closed form (one coating event + one conductor event, no random walk), four
// cases by (coating smooth?, conductor smooth?); operand order follows it so the stochastic branches match the oracle.
#pragma once

struct CCParams { float ieta, iax, iay, cax, cay, thickness; Spec ce, ck, albedo; bool has_medium; };

HK_DEV Spec sp_sqrt(Spec a) { return sp4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w)); }

// :2891-2951 / :3263-3303
HK_DEV CCParams cc_params(const MatCtx& C, const HkMaterial& m, float4 lam, bool regularize) {
    CCParams P;
    P.ieta = m.f[3] == 0.0f ? 1.0f : m.f[3];
    const bool remap = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) != 0;
    P.iax = remap ? sqrtf(m.f[0]) : m.f[0]; P.iay = remap ? sqrtf(m.f[1]) : m.f[1];
    P.cax = remap ? sqrtf(m.f[5]) : m.f[5]; P.cay = remap ? sqrtf(m.f[6]) : m.f[6];
    if (regularize) { P.iax = regularize_alpha(P.iax); P.iay = regularize_alpha(P.iay); P.cax = regularize_alpha(P.cax); P.cay = regularize_alpha(P.cay); }
    if (m.flags & HK_MATFLAG_USE_ETA_K) { P.ce = ior_spectrum(C, m, 0, lam); P.ck = ior_spectrum(C, m, 1, lam); }
    else {   // reflectance mode: eta = 1, k = 2 sqrt(r) / sqrt(1 - r); the cached uplift is of the clamped reflectance
        const Spec r = mat_spec(C, m, 0, lam);
        P.ce = sp(1.0f);
        P.ck = 2.0f * sp_sqrt(r) / sp_sqrt(sp_max0(sp(1.0f) - r) + sp(1.0e-6f));
    }
    P.ce = P.ce / P.ieta; P.ck = P.ck / P.ieta;
    P.thickness = fmaxf(m.f[2], 1.1920929e-7f);
    P.has_medium = !(m.rgb2[0] == 0.0f && m.rgb2[1] == 0.0f && m.rgb2[2] == 0.0f);
    P.albedo = P.has_medium ? pre_bounded(make_pre_bounded(C.T, m.rgb2[0], m.rgb2[1], m.rgb2[2]), lam) : sp(0.0f);   // only read when has_medium
    return P;
}
HK_DEV Spec cc_layer(const CCParams& P, float tr_a, float tr_b) { return P.has_medium ? (tr_a * tr_b) * P.albedo : sp(1.0f); }

HK_DEV BsdfSample sample_coated_conductor(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 n, float4 lam, float2 su, float uc, bool regularize) {
    const float wn = dot3(wo_w, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    const CCParams P = cc_params(C, m, lam, regularize);
    const float ieta = P.ieta;
    const Frame fr = make_frame(n);
    float3 wo = f3(dot3(wo_w, fr.t), dot3(wo_w, fr.b), wn);
    const bool flip = wo.z < 0.0f;
    if (flip) wo = -wo;
    const float cos_o = fabsf(wo.z);
    const bool i_smooth = tr_smooth(P.iax, P.iay), c_smooth = tr_smooth(P.cax, P.cay);
    const float3 mirror = f3(-wo.x, -wo.y, wo.z);
#define HK_CC_WORLD(wl) norm3(to_world(fr, flip ? -(wl) : (wl)))

    if (i_smooth) {
        const float F_i = fresnel_dielectric(cos_o, ieta);
        if (uc < F_i) return bsdf_make(HK_CC_WORLD(mirror), sp(1.0f), 1.0f, true, 1.0f);
        const float s2t = fmaxf(0.0f, 1.0f - cos_o * cos_o) / (ieta * ieta);
        if (s2t >= 1.0f) return bsdf_none();
        const float cos_t_in = sqrtf(1.0f - s2t);
        if (c_smooth) {
            const float3 wb = norm3(f3(-wo.x / ieta, -wo.y / ieta, cos_t_in));
            const Spec F_c = fr_complex4(cos_t_in, P.ce, P.ck);
            const float s2o = fmaxf(0.0f, 1.0f - wb.z * wb.z) * (ieta * ieta);
            if (s2o >= 1.0f) return bsdf_none();
            const float cos_out = sqrtf(1.0f - s2o);
            const float T_in = 1.0f - F_i, T_out = 1.0f - fresnel_dielectric(cos_out, ieta);
            const float tr = P.has_medium ? layer_tr(P.thickness, f3(0, 0, cos_t_in)) : 1.0f;
            return bsdf_make(HK_CC_WORLD(mirror), F_c * T_in * T_out * cc_layer(P, tr, tr) / cos_o, 1.0f - F_i, true, 1.0f);
        }
        const float3 wc = norm3(f3(wo.x / ieta, wo.y / ieta, cos_t_in));
        const float cax = fmaxf(P.cax, 1.0e-4f), cay = fmaxf(P.cay, 1.0e-4f);
        const float3 wm = tr_sample_wm(wc, su, cax, cay);
        const float com = dot3(wc, wm);
        if (com < 0.0f) return bsdf_none();
        const float3 wic = -wc + 2.0f * com * wm;
        if (wic.z < 0.0f) return bsdf_none();
        const Spec F_c = fr_complex4(fabsf(com), P.ce, P.ck);
        const Spec f_c = tr_d(wm, cax, cay) * F_c * tr_g(wc, wic, cax, cay) / (4.0f * fabsf(wc.z) * fabsf(wic.z));
        const float s2o = (wic.x * wic.x + wic.y * wic.y) * (ieta * ieta);
        if (s2o >= 1.0f) return bsdf_none();
        const float cos_out = sqrtf(1.0f - s2o);
        const float T_in = 1.0f - F_i, T_out = 1.0f - fresnel_dielectric(cos_out, ieta);
        const Spec ltr = P.has_medium ? cc_layer(P, layer_tr(P.thickness, f3(0, 0, cos_t_in)), layer_tr(P.thickness, f3(0, 0, wic.z))) : sp(1.0f);
        const float3 wil = norm3(f3(wic.x * ieta, wic.y * ieta, cos_out));
        const float pdf_c = tr_pdf(wc, wm, cax, cay) / (4.0f * fabsf(com));
        return bsdf_make(HK_CC_WORLD(wil), f_c * T_in * T_out * ltr, (1.0f - F_i) * pdf_c, false, 1.0f);
    }

    const float iax = fmaxf(P.iax, 1.0e-4f), iay = fmaxf(P.iay, 1.0e-4f);
    const float3 wm = tr_sample_wm(wo, su, iax, iay);
    const float com = dot3(wo, wm);
    if (com < 0.0f) return bsdf_none();
    const float F_i = fresnel_dielectric(com, ieta);
    if (uc < F_i) {
        float3 wil = -wo + 2.0f * com * wm;
        if (wil.z * wo.z < 0.0f) return bsdf_none();
        const float3 wi = HK_CC_WORLD(wil);
        if (flip) wil = -wil;
        const float D = tr_d(wm, iax, iay), G = tr_g(wo, wil, iax, iay);
        const float pdf = F_i * tr_pdf(wo, wm, iax, iay) / (4.0f * fabsf(com));
        return bsdf_make(wi, sp(D * G / (4.0f * fabsf(wil.z) * cos_o)), pdf, false, 1.0f);
    }
    const float T_in = 1.0f - F_i;
    if (c_smooth) {
        const float cos_b = fabsf(mirror.z);
        const Spec F_c = fr_complex4(cos_b, P.ce, P.ck);
        const float T_out = 1.0f - fresnel_dielectric(cos_b, ieta);
        const float tr = P.has_medium ? layer_tr(P.thickness, mirror) : 1.0f;
        const float pdf = (1.0f - F_i) * tr_pdf(wo, wm, iax, iay) / (4.0f * fabsf(com));
        return bsdf_make(HK_CC_WORLD(mirror), F_c * T_in * T_out * cc_layer(P, tr, tr) / cos_o, pdf, false, 1.0f);
    }
    const float cax = fmaxf(P.cax, 1.0e-4f), cay = fmaxf(P.cay, 1.0e-4f);
    const float3 wmc = tr_sample_wm(wo, su, cax, cay);
    const float comc = dot3(wo, wmc);
    if (comc < 0.0f) return bsdf_none();
    const float3 wil = -wo + 2.0f * comc * wmc;
    if (wil.z * wo.z < 0.0f) return bsdf_none();
    const Spec F_c = fr_complex4(fabsf(comc), P.ce, P.ck);
    const float cos_i = fabsf(wil.z);
    const Spec f_c = tr_d(wmc, cax, cay) * F_c * tr_g(wo, wil, cax, cay) / (4.0f * cos_i * cos_o);
    const float T_out = 1.0f - fresnel_dielectric(cos_i, ieta);
    const Spec ltr = P.has_medium ? cc_layer(P, layer_tr(P.thickness, f3(0, 0, cos_o)), layer_tr(P.thickness, wil)) : sp(1.0f);
    const float pdf = (1.0f - F_i) * tr_pdf(wo, wmc, cax, cay) / (4.0f * fabsf(comc));
    return bsdf_make(HK_CC_WORLD(wil), f_c * T_in * T_out * ltr, pdf, false, 1.0f);
#undef HK_CC_WORLD
}

HK_DEV BsdfEval eval_coated_conductor(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 wi_w, float3 n, float4 lam) {
    const float ci = dot3(wi_w, n), co = dot3(wo_w, n);
    if (ci * co < 0.0f) return eval_none();
    if (fabsf(ci) < 1.0e-6f || fabsf(co) < 1.0e-6f) return eval_none();
    const CCParams P = cc_params(C, m, lam, false);
    const float ieta = P.ieta;
    const Frame fr = make_frame(n);
    float3 wo = f3(dot3(wo_w, fr.t), dot3(wo_w, fr.b), co), wi = f3(dot3(wi_w, fr.t), dot3(wi_w, fr.b), ci);
    if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
    const bool i_smooth = tr_smooth(P.iax, P.iay), c_smooth = tr_smooth(P.cax, P.cay);
    if (i_smooth && c_smooth) return eval_none();
    float3 wh = norm3(wo + wi);
    if (wh.z < 0.0f) wh = -wh;
    const float coh = dot3(wo, wh);
    const float F_wh = fresnel_dielectric(fabsf(coh), ieta);
    const float F_o = fresnel_dielectric(fabsf(wo.z), ieta);
    const float T_o = 1.0f - F_o, T_i = 1.0f - fresnel_dielectric(fabsf(wi.z), ieta);
    const float tr = P.has_medium ? layer_tr(P.thickness, wi) : 1.0f;
    const Spec ltr = cc_layer(P, tr, tr);
    const float denom = 4.0f * fabsf(wi.z) * fabsf(wo.z);
    if (i_smooth) {
        const float cax = fmaxf(P.cax, 1.0e-4f), cay = fmaxf(P.cay, 1.0e-4f);
        const Spec f_c = tr_d(wh, cax, cay) * fr_complex4(fabsf(coh), P.ce, P.ck) * tr_g(wo, wi, cax, cay) / denom;
        return eval_make(f_c * T_o * T_i * ltr, T_o * tr_pdf(wo, wh, cax, cay) / (4.0f * fabsf(coh)));
    }
    const float iax = fmaxf(P.iax, 1.0e-4f), iay = fmaxf(P.iay, 1.0e-4f);
    const float f_interface = tr_d(wh, iax, iay) * F_wh * tr_g(wo, wi, iax, iay) / denom;
    Spec f_c; float pdf_c;
    if (c_smooth) { f_c = fr_complex4(fabsf(wo.z), P.ce, P.ck) / fabsf(wo.z); pdf_c = 1.0f; }
    else {
        const float cax = fmaxf(P.cax, 1.0e-4f), cay = fmaxf(P.cay, 1.0e-4f);
        f_c = tr_d(wh, cax, cay) * fr_complex4(fabsf(coh), P.ce, P.ck) * tr_g(wo, wi, cax, cay) / denom;
        pdf_c = tr_pdf(wo, wh, cax, cay) / (4.0f * fabsf(coh));
    }
    const float pdf_i = F_o * tr_pdf(wo, wh, iax, iay) / (4.0f * fabsf(coh));
    return eval_make(sp(f_interface) + f_c * T_o * T_i * ltr, pdf_i + T_o * pdf_c);
}
