// ok_bsdf_coated_difftrans.h — CPU restatement of the reference's CoatedDiffuseTransmissionMaterial (TEST INFRASTRUCTURE ONLY).
// Follows src/materials/spectral-eval.jl:2249-2338 (diffuse-transmission bottom layer), :2340-2494 (sample),
// :2498-2763 (eval), :2767-2840 (pdf estimate); parameters as in src/materials/coated-diffuse-transmission.jl.
// The reference's functions are its CoatedDiffuse ones (ok_bsdf_layered.h) with the bottom interface swapped for a
// reflect-or-transmit Lambertian, which takes one more random number per sample (uc) — the draw order below is the reference's.
#pragma once
// included from ok_bsdf.h after ok_bsdf_layered.h

namespace ok {

struct DTBottom { Spec refl, trans; float pr_max, pt_max; };

// :2249-2288
inline LSample sample_dt_bottom(V3 wo, V2 u, float uc, const DTBottom& B, uint8_t flags) {
    const float pr = (flags & BXDF_REFLECTION) ? B.pr_max : 0.0f, pt = (flags & BXDF_TRANSMISSION) ? B.pt_max : 0.0f;
    if (pr + pt < 1.0e-10f) return LSample();
    const float prob_reflect = pr / (pr + pt);
    V3 wi = cosine_sample_hemisphere(u);
    if (uc < prob_reflect) {
        if (wo.z < 0.0f) wi = V3(wi.x, wi.y, -wi.z);
        const float ci = std::fabs(wi.z);
        if (ci < 1.0e-6f) return LSample();
        return LSample(B.refl * (1.0f / PI_F), wi, prob_reflect * ci / PI_F, true, false, 1.0f, true);
    }
    if (wo.z > 0.0f) wi = V3(wi.x, wi.y, -wi.z);
    const float ci = std::fabs(wi.z);
    if (ci < 1.0e-6f) return LSample();
    return LSample(B.trans * (1.0f / PI_F), wi, (1.0f - prob_reflect) * ci / PI_F, false, false, 1.0f, true);
}
// :2290-2311
inline Spec eval_dt_bottom(V3 wo, V3 wi, const DTBottom& B, float* pdf = nullptr) {
    if (B.pr_max + B.pt_max < 1.0e-10f) { if (pdf) *pdf = 0.0f; return Spec(); }
    const float aci = std::fabs(wi.z);
    if (same_hemisphere(wo, wi)) {
        if (pdf) *pdf = B.pr_max / (B.pr_max + B.pt_max) * aci / PI_F;
        return B.refl * (1.0f / PI_F);
    }
    if (pdf) *pdf = B.pt_max / (B.pr_max + B.pt_max) * aci / PI_F;
    return B.trans * (1.0f / PI_F);
}
// :2313-2331
inline float pdf_dt_bottom(V3 wo, V3 wi, const DTBottom& B, uint8_t flags = BXDF_ALL) {
    const float pr = (flags & BXDF_REFLECTION) ? B.pr_max : 0.0f, pt = (flags & BXDF_TRANSMISSION) ? B.pt_max : 0.0f;
    if (pr + pt < 1.0e-10f) return 0.0f;
    const float aci = std::fabs(wi.z);
    return same_hemisphere(wo, wi) ? (pr / (pr + pt)) * aci / PI_F : (pt / (pr + pt)) * aci / PI_F;
}

// the parameter block of all three functions (:2350-2376): CoatedDiffuse's, plus the clamped reflectance / transmittance
inline DTBottom dt_bottom_params(const MatCtx& C, const HkMaterial& m, const Wavelengths& l) {
    float r[3], t[3];
    clamp_rgb01(m.rgb0, r); clamp_rgb01(m.rgb2, t);
    DTBottom B;
    B.refl = uplift_rgb(*C.T, r, l); B.trans = uplift_rgb(*C.T, t, l);
    B.pr_max = std::max(std::max(r[0], r[1]), r[2]); B.pt_max = std::max(std::max(t[0], t[1]), t[2]);
    return B;
}

// spectral-eval.jl:2340-2494
inline BSDFSample sample_coated_diffuse_transmission(const MatCtx& C, const HkMaterial& m, V3 wo, V3 n, const Wavelengths& l, V2 sample_u, float rng_in, bool regularize) {
    float wo_dot_n = dot(wo, n);
    if (std::fabs(wo_dot_n) < 1.0e-6f) return BSDFSample();
    CoatedParams P = coated_params(m, regularize);
    DTBottom B = dt_bottom_params(C, m, l);
    Spec albedo = uplift_rgb(*C.T, P.albedo_rgb, l);
    V3 tg, bt; coordinate_system(n, tg, bt);
    V3 wo_l(dot(wo, tg), dot(wo, bt), wo_dot_n);
    bool flip = wo_l.z < 0.0f;
    if (flip) wo_l = -wo_l;
    const float thickness = P.thickness;
    LSample bs = sample_dielectric_interface(wo_l, rng_in, sample_u, P.ax, P.ay, P.eta, BXDF_ALL);
    if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) return BSDFSample();
    if (bs.is_reflection) {
        V3 wl = flip ? -bs.wi : bs.wi;
        return BSDFSample(normalize(tg * wl.x + bt * wl.y + n * wl.z), bs.f, bs.pdf, bs.is_specular, 1.0f);
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
        float uc = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
        bool at_bottom = z == 0.0f;
        LSample bi = at_bottom ? sample_dt_bottom(-w, V2(u1, u2), uc, B, BXDF_ALL)
                               : sample_dielectric_interface(-w, uc, V2(u1, u2), P.ax, P.ay, P.eta, BXDF_ALL);
        if (!bi.valid || bi.pdf == 0.0f || bi.wi.z == 0.0f) return BSDFSample();
        f = f * bi.f;
        pdf *= bi.pdf;
        specular_path = specular_path && bi.is_specular;
        w = bi.wi;
        if (!bi.is_reflection) {     // left the layer: through the coating (top) or through the diffuse base (bottom)
            V3 wl = flip ? -w : w;
            return BSDFSample(normalize(tg * wl.x + bt * wl.y + n * wl.z), f, pdf, specular_path, bi.eta);
        }
        f = f * std::fabs(bi.wi.z);
    }
    return BSDFSample();
}

// spectral-eval.jl:2767-2840
inline float pdf_layered_bsdf_dt(V3 wo, V3 wi, float ax, float ay, float eta, int n_samples, const DTBottom& B) {
    PCG32 rng = pcg32_init(pbrt_hash((uint64_t)0, wi), pbrt_hash(wo));
    bool same_hemi = same_hemisphere(wo, wi);
    bool smooth = tr_effectively_smooth(ax, ay);
    float pdf_sum = 0.0f;
    if (same_hemi && !smooth) pdf_sum += (float)n_samples * pdf_dielectric_interface(wo, wi, ax, ay, eta, BXDF_REFLECTION);
    for (int s = 0; s < n_samples; s++) {
        if (same_hemi) {
            float uc1 = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
            LSample wos = sample_dielectric_interface(wo, uc1, V2(u1, u2), ax, ay, eta, BXDF_TRANSMISSION);
            float uc2 = pcg32_f32(rng), u3 = pcg32_f32(rng), u4 = pcg32_f32(rng);
            LSample wis = sample_dielectric_interface(wi, uc2, V2(u3, u4), ax, ay, eta, BXDF_TRANSMISSION);
            if (wos.valid && wos.pdf > 0.0f && wis.valid && wis.pdf > 0.0f) {
                if (smooth) pdf_sum += pdf_dt_bottom(-wos.wi, -wis.wi, B);
                else {
                    float u5 = pcg32_f32(rng), u6 = pcg32_f32(rng), uc3 = pcg32_f32(rng);
                    LSample rs = sample_dt_bottom(-wos.wi, V2(u5, u6), uc3, B, BXDF_ALL);
                    if (rs.valid && rs.pdf > 0.0f) {
                        float r_pdf = pdf_dt_bottom(-wos.wi, -wis.wi, B);
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
            float uc2 = pcg32_f32(rng), u3 = pcg32_f32(rng), u4 = pcg32_f32(rng);
            LSample wis = sample_dt_bottom(wi, V2(u3, u4), uc2, B, BXDF_TRANSMISSION);
            if (!wis.valid || wis.pdf == 0.0f || wis.is_reflection) continue;
            if (smooth) pdf_sum += pdf_dt_bottom(-wos.wi, wi, B);
            else pdf_sum += (pdf_dielectric_interface(wo, -wis.wi, ax, ay, eta) + pdf_dt_bottom(-wos.wi, wi, B)) / 2.0f;
        }
    }
    return lerpf(0.9f, 1.0f / (4.0f * PI_F), pdf_sum / (float)n_samples);
}

// spectral-eval.jl:2498-2763
inline BSDFEval eval_coated_diffuse_transmission(const MatCtx& C, const HkMaterial& m, V3 wo, V3 wi, V3 n, const Wavelengths& l) {
    CoatedParams P = coated_params(m, false);
    DTBottom B = dt_bottom_params(C, m, l);
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
        LSample wis = exit_at_bottom ? sample_dt_bottom(wi_l, V2(u1, u2), uc, B, BXDF_TRANSMISSION)
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
                    if (exit_at_bottom) wt = power_heuristic(1, wis.pdf, 1, hg_phase_pdf(g, dot(-w, -wis.wi)));
                    else wt = !smooth ? power_heuristic(1, wis.pdf, 1, hg_phase_pdf(g, dot(-w, -wis.wi))) : 1.0f;
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
                        if (exit_at_bottom) fe2 = eval_dt_bottom(-w, wi_l, B, &exit_pdf);
                        else {
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
                LSample bs = exit_at_bottom ? sample_dt_bottom(-w, V2(v1, v2), uc2, B, BXDF_REFLECTION)
                                            : sample_dielectric_interface(-w, uc2, V2(v1, v2), ax, ay, eta, BXDF_REFLECTION);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * std::fabs(bs.wi.z) / bs.pdf;
                w = bs.wi;
            } else {
                bool ne_bottom = z == 0.0f;
                bool non_exit_specular = !ne_bottom && smooth;
                if (!non_exit_specular) {
                    Spec f_nee = ne_bottom ? eval_dt_bottom(-w, -wis.wi, B) : eval_dielectric_interface(-w, -wis.wi, ax, ay, eta);
                    if (max_component(f_nee) > 0.0f) {
                        float wt = 1.0f;
                        if (!exit_at_bottom || !smooth) {
                            float nee_pdf = ne_bottom ? pdf_dt_bottom(-w, -wis.wi, B) : pdf_dielectric_interface(-w, -wis.wi, ax, ay, eta);
                            wt = power_heuristic(1, wis.pdf, 1, nee_pdf);
                        }
                        fr = fr + beta * f_nee * std::fabs(wis.wi.z) * wt * layer_transmittance(thickness, wis.wi) * wis.f / wis.pdf;
                    }
                }
                float uc2 = pcg32_f32(rng), v1 = pcg32_f32(rng), v2 = pcg32_f32(rng);
                LSample bs = ne_bottom ? sample_dt_bottom(-w, V2(v1, v2), uc2, B, BXDF_REFLECTION)
                                       : sample_dielectric_interface(-w, uc2, V2(v1, v2), ax, ay, eta, BXDF_REFLECTION);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * std::fabs(bs.wi.z) / bs.pdf;
                w = bs.wi;
                if (!smooth || exit_at_bottom) {
                    Spec fe3 = exit_at_bottom ? eval_dt_bottom(-w, wi_l, B) : eval_dielectric_interface(-w, wi_l, ax, ay, eta);
                    if (max_component(fe3) > 0.0f) {
                        float wt3 = 1.0f;
                        if (!non_exit_specular) {
                            float ep3 = exit_at_bottom ? pdf_dt_bottom(-w, wi_l, B) : pdf_dielectric_interface(-w, wi_l, ax, ay, eta, BXDF_TRANSMISSION);
                            wt3 = power_heuristic(1, bs.pdf, 1, ep3);
                        }
                        fr = fr + beta * layer_transmittance(thickness, bs.wi) * fe3 * wt3;
                    }
                }
            }
        }
    }
    fr = fr / (float)P.n_samples;
    return BSDFEval(fr, pdf_layered_bsdf_dt(wo_l, wi_l, ax, ay, eta, P.n_samples, B));
}

}  // namespace ok
