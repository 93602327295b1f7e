// hk_bsdf_coated_difftrans.cuh — CoatedDiffuseTransmissionMaterial: dielectric coating over a reflect-or-transmit Lambertian.
// Reference: src/materials/spectral-eval.jl:2249-2338 (bottom layer), :2340-2494 (sample), :2498-2763 (eval),
// :2767-2840 (pdf estimate); src/materials/coated-diffuse-transmission.jl.  Same LayeredBxDF walk as CoatedDiffuse
// (hk_bsdf_layered.cuh) with the base swapped; the base takes one more random number per sample, drawn in the reference's order.
#pragma once

struct DTBase { Spec refl, trans; float pr_max, pt_max; };

HK_DEV DTBase dt_base(const MatCtx& C, const HkMaterial& m, float4 lam) {      // :2350-2376
    DTBase B;
    B.refl = mat_spec(C, m, 0, lam);                                            // rgb_to_spectrum clamps to [0, 1] itself
    B.trans = pre_bounded(make_pre_bounded(C.T, m.rgb2[0], m.rgb2[1], m.rgb2[2]), lam);
    B.pr_max = fmaxf(fmaxf(clampf(m.rgb0[0], 0.0f, 1.0f), clampf(m.rgb0[1], 0.0f, 1.0f)), clampf(m.rgb0[2], 0.0f, 1.0f));
    B.pt_max = fmaxf(fmaxf(clampf(m.rgb2[0], 0.0f, 1.0f), clampf(m.rgb2[1], 0.0f, 1.0f)), clampf(m.rgb2[2], 0.0f, 1.0f));
    return B;
}
HK_DEV IfaceSample dt_sample(float3 wo, float2 u, float uc, const DTBase& B, uint32_t flags) {      // :2249-2288
    const float pr = (flags & HK_REFL) ? B.pr_max : 0.0f, pt = (flags & HK_TRANS) ? B.pt_max : 0.0f;
    if (pr + pt < 1.0e-10f) return iface_invalid();
    const float prob = pr / (pr + pt);
    float3 wi = cosine_sample_hemisphere(u);
    const bool refl = uc < prob;
    if (refl ? (wo.z < 0.0f) : (wo.z > 0.0f)) wi.z = -wi.z;
    const float c = fabsf(wi.z);
    if (c < 1.0e-6f) return iface_invalid();
    if (refl) return iface_make(B.refl * (1.0f / HK_PI), wi, prob * c / HK_PI, true, false, 1.0f);
    return iface_make(B.trans * (1.0f / HK_PI), wi, (1.0f - prob) * c / HK_PI, false, false, 1.0f);
}
HK_DEV float dt_pdf(float3 wo, float3 wi, const DTBase& B, uint32_t flags) {                      // :2313-2331
    const float pr = (flags & HK_REFL) ? B.pr_max : 0.0f, pt = (flags & HK_TRANS) ? B.pt_max : 0.0f;
    if (pr + pt < 1.0e-10f) return 0.0f;
    const float c = fabsf(wi.z);
    return same_hemi(wo, wi) ? (pr / (pr + pt)) * c / HK_PI : (pt / (pr + pt)) * c / HK_PI;
}
HK_DEV Spec dt_eval(float3 wo, float3 wi, const DTBase& B, float& pdf) {                          // :2290-2311
    if (B.pr_max + B.pt_max < 1.0e-10f) { pdf = 0.0f; return sp(0.0f); }
    const float c = fabsf(wi.z);
    if (same_hemi(wo, wi)) { pdf = B.pr_max / (B.pr_max + B.pt_max) * c / HK_PI; return B.refl * (1.0f / HK_PI); }
    pdf = B.pt_max / (B.pr_max + B.pt_max) * c / HK_PI;
    return B.trans * (1.0f / HK_PI);
}
HK_DEV Spec dt_eval(float3 wo, float3 wi, const DTBase& B) { float p; return dt_eval(wo, wi, B, p); }

// sample: :2340-2494
HK_DEV BsdfSample sample_coated_difftrans(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam, float2 su, float uc_in, bool regularize) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    CoatParams P = coat_params(m, regularize);
    const DTBase B = dt_base(C, m, lam);
    Spec albedo = mat_spec(C, m, 1, lam);
    Frame fr = make_frame(n);
    float3 wl = f3(dot3(wo, fr.t), dot3(wo, fr.b), wn);
    const bool flip = wl.z < 0.0f;
    if (flip) wl = -wl;
    IfaceSample bs = coat_sample(wl, uc_in, su, P.ax, P.ay, P.eta, HK_RT_ALL);
    if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) return bsdf_none();
    if (bs.reflection) {
        float3 o = flip ? -bs.wi : bs.wi;
        return bsdf_make(norm3(to_world(fr, o)), bs.f, bs.pdf, bs.specular, 1.0f);
    }
    float3 w = bs.wi;
    bool spec_path = bs.specular;
    Spec f = bs.f * fabsf(w.z);
    float pdf = bs.pdf, z = P.thickness;
    Pcg32 rng = pcg32_init(hash_u64_f3(0ull, wl), hash_f_f2(uc_in, su.x, su.y));
    for (int depth = 0; depth < P.max_depth; depth++) {
        float rrb = sp_maxc(f) / pdf;
        if (depth > 3 && rrb < 0.25f) {
            float q = fmaxf(0.0f, 1.0f - rrb);
            if (pcg32_f32(rng) < q) return bsdf_none();
            pdf *= 1.0f - q;
        }
        if (w.z == 0.0f) return bsdf_none();
        if (P.has_medium) {
            float dz = -dm_logf(1.0f - pcg32_f32(rng)) / (1.0f / fabsf(w.z));
            float zp = w.z > 0.0f ? z + dz : z - dz;
            if (zp == z) return bsdf_none();
            if (0.0f < zp && zp < P.thickness) {
                float p1 = pcg32_f32(rng), p2 = pcg32_f32(rng), pp;
                float3 wp = hg_sample_layer(P.g, -w, make_float2(p1, p2), pp);
                if (pp == 0.0f || wp.z == 0.0f) return bsdf_none();
                f = f * albedo * pp; pdf *= pp; spec_path = false; w = wp; z = zp;
                continue;
            }
            z = clampf(zp, 0.0f, P.thickness);
        } else {
            z = (z == P.thickness) ? 0.0f : P.thickness;
            f = f * layer_tr(P.thickness, w);
        }
        float uc = pcg32_f32(rng), u1 = pcg32_f32(rng), u2 = pcg32_f32(rng);
        IfaceSample bi = (z == 0.0f) ? dt_sample(-w, make_float2(u1, u2), uc, B, HK_RT_ALL)
                                     : coat_sample(-w, uc, make_float2(u1, u2), P.ax, P.ay, P.eta, HK_RT_ALL);
        if (!bi.valid || bi.pdf == 0.0f || bi.wi.z == 0.0f) return bsdf_none();
        f = f * bi.f; pdf *= bi.pdf; spec_path = spec_path && bi.specular; w = bi.wi;
        if (!bi.reflection) {      // left the layer: through the coating (top) or through the diffuse base (bottom)
            float3 o = flip ? -w : w;
            return bsdf_make(norm3(to_world(fr, o)), f, pdf, spec_path, bi.eta);
        }
        f = f * fabsf(bi.wi.z);
    }
    return bsdf_none();
}

// pdf estimate: :2767-2840
HK_DEV float coated_difftrans_pdf(float3 wo, float3 wi, const CoatParams& P, const DTBase& B) {
    Pcg32 rng = pcg32_init(hash_u64_f3(0ull, wi), hash_f3(wo));
    const bool sh = same_hemi(wo, wi), smooth = tr_smooth(P.ax, P.ay);
    float sum = 0.0f;
    if (sh && !smooth) sum += (float)P.n_samples * coat_pdf(wo, wi, P.ax, P.ay, P.eta, HK_REFL);
    for (int s = 0; s < P.n_samples; s++) {
        if (sh) {
            float a0 = pcg32_f32(rng), a1 = pcg32_f32(rng), a2 = pcg32_f32(rng);
            IfaceSample wos = coat_sample(wo, a0, make_float2(a1, a2), P.ax, P.ay, P.eta, HK_TRANS);
            float b0 = pcg32_f32(rng), b1 = pcg32_f32(rng), b2 = pcg32_f32(rng);
            IfaceSample wis = coat_sample(wi, b0, make_float2(b1, b2), P.ax, P.ay, P.eta, HK_TRANS);
            if (wos.valid && wos.pdf > 0.0f && wis.valid && wis.pdf > 0.0f) {
                if (smooth) sum += dt_pdf(-wos.wi, -wis.wi, B, HK_RT_ALL);
                else {
                    float c1 = pcg32_f32(rng), c2 = pcg32_f32(rng), c0 = pcg32_f32(rng);
                    IfaceSample rs = dt_sample(-wos.wi, make_float2(c1, c2), c0, B, HK_RT_ALL);
                    if (rs.valid && rs.pdf > 0.0f) {
                        float rp = dt_pdf(-wos.wi, -wis.wi, B, HK_RT_ALL);
                        sum += power_heur(wis.pdf, rp) * rp;
                        float tp = coat_pdf(-rs.wi, wi, P.ax, P.ay, P.eta, HK_RT_ALL);
                        sum += power_heur(rs.pdf, tp) * tp;
                    }
                }
            }
        } else {
            float a0 = pcg32_f32(rng), a1 = pcg32_f32(rng), a2 = pcg32_f32(rng);
            IfaceSample wos = coat_sample(wo, a0, make_float2(a1, a2), P.ax, P.ay, P.eta, HK_TRANS);
            if (!wos.valid || wos.pdf == 0.0f || wos.reflection) continue;
            float b0 = pcg32_f32(rng), b1 = pcg32_f32(rng), b2 = pcg32_f32(rng);
            IfaceSample wis = dt_sample(wi, make_float2(b1, b2), b0, B, HK_TRANS);
            if (!wis.valid || wis.pdf == 0.0f || wis.reflection) continue;
            if (smooth) sum += dt_pdf(-wos.wi, wi, B, HK_RT_ALL);
            else sum += (coat_pdf(wo, -wis.wi, P.ax, P.ay, P.eta, HK_RT_ALL) + dt_pdf(-wos.wi, wi, B, HK_RT_ALL)) / 2.0f;
        }
    }
    return lerpf(0.9f, 1.0f / (4.0f * HK_PI), sum / (float)P.n_samples);
}

// eval: :2498-2763
HK_DEV BsdfEval eval_coated_difftrans(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 wi_w, float3 n, float4 lam) {
    CoatParams P = coat_params(m, false);
    const DTBase B = dt_base(C, m, lam);
    Spec albedo = mat_spec(C, m, 1, lam);
    const float th = P.thickness, ax = P.ax, ay = P.ay, eta = P.eta, g = P.g;
    Frame fr = make_frame(n);
    float3 wo = f3(dot3(wo_w, fr.t), dot3(wo_w, fr.b), dot3(wo_w, n));
    float3 wi = f3(dot3(wi_w, fr.t), dot3(wi_w, fr.b), dot3(wi_w, n));
    if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
    if (fabsf(wo.z) < 1.0e-6f || fabsf(wi.z) < 1.0e-6f) return eval_none();
    const bool sh = same_hemi(wo, wi);
    const bool exit_bottom = !sh;                 // same_hemi XOR entered_top(=true)
    const float exit_z = exit_bottom ? 0.0f : th;
    const bool smooth = tr_smooth(ax, ay);
    Spec acc = sp(0.0f);
    if (sh) acc = acc + coat_eval(wo, wi, ax, ay, eta) * (float)P.n_samples;
    Pcg32 rng = pcg32_init(hash_u64_f3(0ull, wo), hash_f3(wi));
    for (int s = 0; s < P.n_samples; s++) {
        float a0 = pcg32_f32(rng), a1 = pcg32_f32(rng), a2 = pcg32_f32(rng);
        IfaceSample wos = coat_sample(wo, a0, make_float2(a1, a2), ax, ay, eta, HK_TRANS);
        if (!wos.valid || wos.pdf == 0.0f || wos.wi.z == 0.0f) continue;
        float b0 = pcg32_f32(rng), b1 = pcg32_f32(rng), b2 = pcg32_f32(rng);
        IfaceSample wis = exit_bottom ? dt_sample(wi, make_float2(b1, b2), b0, B, HK_TRANS)
                                      : coat_sample(wi, b0, make_float2(b1, b2), ax, ay, eta, HK_TRANS);
        if (!wis.valid || wis.pdf == 0.0f || wis.wi.z == 0.0f) continue;
        Spec beta = wos.f * fabsf(wos.wi.z) / wos.pdf;
        float z = th;
        float3 w = wos.wi;
        for (int depth = 0; depth < P.max_depth; depth++) {
            if (depth > 3 && sp_maxc(beta) < 0.25f) {
                float q = fmaxf(0.0f, 1.0f - sp_maxc(beta));
                if (pcg32_f32(rng) < q) break;
                beta = beta / (1.0f - q);
            }
            if (P.has_medium) {
                float dz = -dm_logf(1.0f - pcg32_f32(rng)) / (1.0f / fabsf(w.z));
                float zp = w.z > 0.0f ? z + dz : z - dz;
                if (zp == z) continue;
                if (0.0f < zp && zp < th) {
                    float ph = hg_phase(g, dot3(-w, -wis.wi));
                    float wt = (exit_bottom || !smooth) ? power_heur(wis.pdf, ph) : 1.0f;
                    acc = acc + beta * albedo * ph * wt * layer_tr(zp - exit_z, wis.wi) * wis.f / wis.pdf;
                    float p1 = pcg32_f32(rng), p2 = pcg32_f32(rng), pp;
                    float3 wp = hg_sample_layer(g, -w, make_float2(p1, p2), pp);
                    if (pp == 0.0f || wp.z == 0.0f) break;
                    beta = beta * albedo * pp / pp;
                    w = wp; z = zp;
                    if ((z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f)) {
                        Spec fe; float ep;
                        if (exit_bottom) fe = dt_eval(-w, wi, B, ep);
                        else if (!smooth) { fe = coat_eval(-w, wi, ax, ay, eta); ep = coat_pdf(-w, wi, ax, ay, eta, HK_TRANS); }
                        else continue;
                        if (sp_maxc(fe) > 0.0f) acc = acc + beta * layer_tr(zp - exit_z, wp) * fe * power_heur(pp, ep);
                    }
                    continue;
                }
                z = clampf(zp, 0.0f, th);
            } else {
                z = (z == th) ? 0.0f : th;
                beta = beta * layer_tr(th, w);
            }
            if (z == exit_z) {
                float c0 = pcg32_f32(rng), c1 = pcg32_f32(rng), c2 = pcg32_f32(rng);
                IfaceSample bs = exit_bottom ? dt_sample(-w, make_float2(c1, c2), c0, B, HK_REFL)
                                             : coat_sample(-w, c0, make_float2(c1, c2), ax, ay, eta, HK_REFL);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * fabsf(bs.wi.z) / bs.pdf;
                w = bs.wi;
            } else {
                const bool ne_bottom = z == 0.0f;
                const bool ne_spec = !ne_bottom && smooth;
                if (!ne_spec) {
                    Spec fn = ne_bottom ? dt_eval(-w, -wis.wi, B) : coat_eval(-w, -wis.wi, ax, ay, eta);
                    if (sp_maxc(fn) > 0.0f) {
                        float wt = 1.0f;
                        if (!exit_bottom || !smooth) {
                            float np = ne_bottom ? dt_pdf(-w, -wis.wi, B, HK_RT_ALL) : coat_pdf(-w, -wis.wi, ax, ay, eta, HK_RT_ALL);
                            wt = power_heur(wis.pdf, np);
                        }
                        acc = acc + beta * fn * fabsf(wis.wi.z) * wt * layer_tr(th, wis.wi) * wis.f / wis.pdf;
                    }
                }
                float c0 = pcg32_f32(rng), c1 = pcg32_f32(rng), c2 = pcg32_f32(rng);
                IfaceSample bs = ne_bottom ? dt_sample(-w, make_float2(c1, c2), c0, B, HK_REFL)
                                           : coat_sample(-w, c0, make_float2(c1, c2), ax, ay, eta, HK_REFL);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * fabsf(bs.wi.z) / bs.pdf;
                w = bs.wi;
                if (!smooth || exit_bottom) {
                    Spec f3e = exit_bottom ? dt_eval(-w, wi, B) : coat_eval(-w, wi, ax, ay, eta);
                    if (sp_maxc(f3e) > 0.0f) {
                        float wt3 = 1.0f;
                        if (!ne_spec) wt3 = power_heur(bs.pdf, exit_bottom ? dt_pdf(-w, wi, B, HK_RT_ALL) : coat_pdf(-w, wi, ax, ay, eta, HK_TRANS));
                        acc = acc + beta * layer_tr(th, bs.wi) * f3e * wt3;
                    }
                }
            }
        }
    }
    acc = acc / (float)P.n_samples;
    return eval_make(acc, coated_difftrans_pdf(wo, wi, P, B));
}
