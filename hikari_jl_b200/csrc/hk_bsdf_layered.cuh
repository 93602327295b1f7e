// hk_bsdf_layered.cuh — CoatedDiffuse: stochastic random walk between a (rough) dielectric coat and a
// Lambertian base (pbrt-v4 LayeredBxDF as ported by src/materials/spectral-eval.jl:823-1937).
// The walk's private RNG is PCG32 seeded from hashes of the directions, so results are a pure function of
// the inputs — which is what lets the GPU and the CPU restatement agree sample for sample.
// Quirks of the reference that are kept on purpose: lerp() argument order at :1936, `phase_p / phase_p`
// at :1712, diffuse exit sampled with BXDF_TRANSMISSION (always invalid) at :1649.
#pragma once

#define HK_REFL 1u
#define HK_TRANS 2u
#define HK_RT_ALL 3u

struct IfaceSample { Spec f; float3 wi; float pdf; float eta; bool reflection, specular, valid; };
HK_DEV IfaceSample iface_invalid() { IfaceSample s; s.f = sp(0.0f); s.wi = f3(0, 0, 0); s.pdf = 0.0f; s.eta = 1.0f; s.reflection = false; s.specular = false; s.valid = false; return s; }
HK_DEV IfaceSample iface_make(Spec f, float3 wi, float pdf, bool refl, bool spec, float eta) { IfaceSample s; s.f = f; s.wi = wi; s.pdf = pdf; s.eta = eta; s.reflection = refl; s.specular = spec; s.valid = true; return s; }

HK_DEV float layer_tr(float thickness, float3 w) { return fabsf(thickness) <= 1.1920929e-7f ? 1.0f : dm_expf(-fabsf(thickness / w.z)); }   // :837-840
HK_DEV float hg_phase(float g, float c) {                                                                                                // :879-883
    float g2 = g * g, d = 1.0f + g2 - 2.0f * g * c;
    return (1.0f - g2) / (4.0f * HK_PI * d * sqrtf(fmaxf(1.0e-10f, d)));
}
HK_NI_LAYERED float3 hg_sample_layer(float g, float3 wo, float2 u, float& p) {                                                                    // :847-872
    float c;
    if (fabsf(g) < 1.0e-3f) c = 1.0f - 2.0f * u.x;
    else { float g2 = g * g; float q = (1.0f - g2) / (1.0f - g + 2.0f * g * u.x); c = clampf((1.0f + g2 - q * q) / (2.0f * g), -1.0f, 1.0f); }
    float s = sqrtf(fmaxf(0.0f, 1.0f - c * c)), phi = 2.0f * HK_PI * u.y;
    Frame fr = make_frame(-wo);
    float3 wi = norm3(s * dm_cosf(phi) * fr.t + s * dm_sinf(phi) * fr.b + c * (-wo));
    p = hg_phase(g, c);
    return wi;
}
HK_DEV bool refract_flat(float3 wo, float eta, float3& wi, float& etap) {                                                                  // :1072-1093
    float ci = wo.z;
    etap = ci > 0.0f ? eta : 1.0f / eta;
    float s2t = fmaxf(0.0f, 1.0f - ci * ci) / (etap * etap);
    if (s2t >= 1.0f) return false;
    float ct = sqrtf(1.0f - s2t);
    wi = norm3(f3(-wo.x / etap, -wo.y / etap, ci > 0.0f ? -ct : ct));
    return true;
}
HK_DEV bool refract_mf(float3 wo, float3 wm, float eta, float3& wi, float& etap) {                                                         // :1100-1120
    float ci = dot3(wo, wm);
    etap = ci > 0.0f ? eta : 1.0f / eta;
    float s2t = fmaxf(0.0f, 1.0f - ci * ci) / (etap * etap);
    if (s2t >= 1.0f) return false;
    float ct = sqrtf(1.0f - s2t);
    wi = norm3(-wo / etap + (ci / etap + (ci > 0.0f ? -ct : ct)) * wm);
    return true;
}
HK_NI_LAYERED IfaceSample coat_sample(float3 wo, float uc, float2 u, float ax, float ay, float eta, uint32_t flags) {                              // :973-1063
    if (tr_smooth(ax, ay) || eta == 1.0f) {
        float R = fresnel_dielectric(wo.z, eta), T = 1.0f - R;
        float pr = (flags & HK_REFL) ? R : 0.0f, pt = (flags & HK_TRANS) ? T : 0.0f;
        if (pr == 0.0f && pt == 0.0f) return iface_invalid();
        if (uc < pr / (pr + pt)) { float3 wi = f3(-wo.x, -wo.y, wo.z); return iface_make(sp(R / fabsf(wi.z)), wi, pr / (pr + pt), true, true, 1.0f); }
        float3 wi; float etap;
        if (!refract_flat(wo, eta, wi, etap)) return iface_invalid();
        return iface_make(sp(T / fabsf(wi.z)), wi, pt / (pr + pt), false, true, etap);
    }
    float3 wm = tr_sample_wm(wo, u, ax, ay);
    float com = dot3(wo, wm);
    float R = fresnel_dielectric(com, eta), T = 1.0f - R;
    float pr = (flags & HK_REFL) ? R : 0.0f, pt = (flags & HK_TRANS) ? T : 0.0f;
    if (pr == 0.0f && pt == 0.0f) return iface_invalid();
    if (uc < pr / (pr + pt)) {
        float3 wi = reflect3(wo, wm);
        if (!same_hemi(wo, wi)) return iface_invalid();
        float pdf = tr_pdf(wo, wm, ax, ay) / (4.0f * fabsf(com)) * pr / (pr + pt);
        float f = tr_d(wm, ax, ay) * tr_g(wo, wi, ax, ay) * R / (4.0f * wo.z * wi.z);
        return iface_make(sp(f), wi, pdf, true, false, 1.0f);
    }
    float3 wi; float etap;
    if (!refract_mf(wo, wm, eta, wi, etap) || same_hemi(wo, wi) || wi.z == 0.0f) return iface_invalid();
    float dd = dot3(wi, wm) + dot3(wo, wm) / etap;
    float den = dd * dd;
    float pdf = tr_pdf(wo, wm, ax, ay) * (fabsf(dot3(wi, wm)) / den) * pt / (pr + pt);
    float f = T * tr_d(wm, ax, ay) * tr_g(wo, wi, ax, ay) * fabsf(dot3(wi, wm) * dot3(wo, wm) / (wi.z * wo.z * den));
    return iface_make(sp(f), wi, pdf, false, false, etap);
}
HK_DEV IfaceSample base_sample(float3 wo, float2 u, Spec refl, uint32_t flags) {                                                          // :1144-1171
    if ((flags & HK_REFL) == 0) return iface_invalid();
    float3 wi = cosine_sample_hemisphere(u);
    if (wo.z < 0.0f) wi.z = -wi.z;
    float c = fabsf(wi.z);
    if (c < 1.0e-6f) return iface_invalid();
    return iface_make(refl * (1.0f / HK_PI), wi, c / HK_PI, true, false, 1.0f);
}
HK_DEV Spec base_eval(float3 wo, float3 wi, Spec refl) { return same_hemi(wo, wi) ? refl * (1.0f / HK_PI) : sp(0.0f); }                    // :1178-1187
HK_DEV float base_pdf(float3 wo, float3 wi) { return same_hemi(wo, wi) ? fabsf(wi.z) / HK_PI : 0.0f; }                                      // :1194-1199
HK_DEV float power_heur(float fp, float gp) { float f2 = fp * fp, g2 = gp * gp; return (f2 + g2 == 0.0f) ? 0.0f : f2 / (f2 + g2); }        // :1206-1215 (nf = ng = 1)
HK_NI_LAYERED Spec coat_eval(float3 wo, float3 wi, float ax, float ay, float eta) {                                                              // :1426-1486
    if (tr_smooth(ax, ay) || eta == 1.0f) return sp(0.0f);
    if (same_hemi(wo, wi)) {
        float3 wh = norm3(wo + wi);
        if (wh.z < 0.0f) wh = -wh;
        float R = fresnel_dielectric(dot3(wo, wh), eta);
        return sp(tr_d(wh, ax, ay) * tr_g(wo, wi, ax, ay) * R / (4.0f * wo.z * wi.z));
    }
    float etap = wo.z > 0.0f ? eta : 1.0f / eta;
    float3 wh = norm3(wo + wi * etap);
    if (wh.z < 0.0f) wh = -wh;
    float coh = dot3(wo, wh), cih = dot3(wi, wh);
    if (coh * cih > 0.0f) return sp(0.0f);
    float T = 1.0f - fresnel_dielectric(coh, eta);
    float dd = cih + coh / etap;
    float den = dd * dd;
    return sp(T * tr_d(wh, ax, ay) * tr_g(wo, wi, ax, ay) * fabsf(cih * coh / (wo.z * wi.z * den)));
}
HK_NI_LAYERED float coat_pdf(float3 wo, float3 wi, float ax, float ay, float eta, uint32_t flags) {                                               // :1493-1554
    if (tr_smooth(ax, ay) || eta == 1.0f) return 0.0f;
    if (same_hemi(wo, wi)) {
        if ((flags & HK_REFL) == 0) return 0.0f;
        float3 wh = norm3(wo + wi);
        if (wh.z < 0.0f) wh = -wh;
        float coh = fabsf(dot3(wo, wh));
        float R = fresnel_dielectric(coh, eta), T = 1.0f - R;
        float pr = R, pt = (flags & HK_TRANS) ? T : 0.0f;
        return tr_pdf(wo, wh, ax, ay) / (4.0f * coh) * pr / (pr + pt);
    }
    if ((flags & HK_TRANS) == 0) return 0.0f;
    float etap = wo.z > 0.0f ? eta : 1.0f / eta;
    float3 wh = norm3(wo + wi * etap);
    if (wh.z < 0.0f) wh = -wh;
    float coh = dot3(wo, wh), cih = dot3(wi, wh);
    if (coh * cih > 0.0f) return 0.0f;
    float R = fresnel_dielectric(fabsf(coh), eta), T = 1.0f - R;
    float pr = (flags & HK_REFL) ? R : 0.0f, pt = T;
    float dd = cih + coh / etap;
    return tr_pdf(wo, wh, ax, ay) * (fabsf(cih) / (dd * dd)) * pt / (pr + pt);
}

struct CoatParams { float eta, thickness, g, ax, ay; int max_depth, n_samples; bool has_medium; };
HK_DEV CoatParams coat_params(const HkMaterial& m, bool regularize) {
    CoatParams p;
    p.eta = m.f[3];
    p.thickness = fmaxf(m.f[2], 1.1920929e-7f);
    p.g = clampf(m.f[4], -0.99f, 0.99f);
    bool remap = (m.flags & HK_MATFLAG_REMAP_ROUGHNESS) != 0;
    p.ax = remap ? sqrtf(m.f[0]) : m.f[0];
    p.ay = remap ? sqrtf(m.f[1]) : m.f[1];
    if (regularize) { p.ax = regularize_alpha(p.ax); p.ay = regularize_alpha(p.ay); }
    p.max_depth = m.ival[0]; p.n_samples = m.ival[1];
    p.has_medium = !(m.rgb1[0] == 0.0f && m.rgb1[1] == 0.0f && m.rgb1[2] == 0.0f);
    return p;
}

// sample: :1232-1418
HK_DEV BsdfSample sample_coated_diffuse(const MatCtx& C, const HkMaterial& m, float3 wo, float3 n, float4 lam, float2 su, float uc_in, bool regularize) {
    float wn = dot3(wo, n);
    if (fabsf(wn) < 1.0e-6f) return bsdf_none();
    CoatParams P = coat_params(m, regularize);
    Spec refl = mat_spec(C, m, 0, lam);
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
        IfaceSample bi = (z == 0.0f) ? base_sample(-w, make_float2(u1, u2), refl, HK_RT_ALL)
                                     : coat_sample(-w, uc, make_float2(u1, u2), P.ax, P.ay, P.eta, HK_RT_ALL);
        if (!bi.valid || bi.pdf == 0.0f || bi.wi.z == 0.0f) return bsdf_none();
        f = f * bi.f; pdf *= bi.pdf; spec_path = spec_path && bi.specular; w = bi.wi;
        if (!bi.reflection) {
            float3 o = flip ? -w : w;
            return bsdf_make(norm3(to_world(fr, o)), f, pdf, spec_path, bi.eta);
        }
        f = f * fabsf(bi.wi.z);
    }
    return bsdf_none();
}

// pdf estimate: :1848-1937
HK_DEV float coated_pdf(float3 wo, float3 wi, const CoatParams& P, Spec refl) {
    Pcg32 rng = pcg32_init(hash_u64_f3(0ull, wi), hash_f3(wo));
    const bool sh = same_hemi(wo, wi), smooth = tr_smooth(P.ax, P.ay);
    float sum = 0.0f;
    if (sh) sum += smooth ? (float)P.n_samples * 0.0f : (float)P.n_samples * coat_pdf(wo, wi, P.ax, P.ay, P.eta, HK_REFL);
    for (int s = 0; s < P.n_samples; s++) {
        if (sh) {
            float a0 = pcg32_f32(rng), a1 = pcg32_f32(rng), a2 = pcg32_f32(rng);
            IfaceSample wos = coat_sample(wo, a0, make_float2(a1, a2), P.ax, P.ay, P.eta, HK_TRANS);
            float b0 = pcg32_f32(rng), b1 = pcg32_f32(rng), b2 = pcg32_f32(rng);
            IfaceSample wis = coat_sample(wi, b0, make_float2(b1, b2), P.ax, P.ay, P.eta, HK_TRANS);
            if (wos.valid && wos.pdf > 0.0f && wis.valid && wis.pdf > 0.0f) {
                if (smooth) sum += base_pdf(-wos.wi, -wis.wi);
                else {
                    float c1 = pcg32_f32(rng), c2 = pcg32_f32(rng);
                    IfaceSample rs = base_sample(-wos.wi, make_float2(c1, c2), refl, HK_RT_ALL);
                    if (rs.valid && rs.pdf > 0.0f) {
                        float rp = base_pdf(-wos.wi, -wis.wi);
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
            float b1 = pcg32_f32(rng), b2 = pcg32_f32(rng);
            IfaceSample wis = base_sample(wi, make_float2(b1, b2), refl, HK_TRANS);
            if (!wis.valid || wis.pdf == 0.0f || wis.reflection) continue;
            if (smooth) sum += base_pdf(-wos.wi, wi);
            else sum += (coat_pdf(wo, -wis.wi, P.ax, P.ay, P.eta, HK_RT_ALL) + base_pdf(-wos.wi, wi)) / 2.0f;
        }
    }
    return lerpf(0.9f, 1.0f / (4.0f * HK_PI), sum / (float)P.n_samples);
}

// eval: :1564-1840
HK_DEV BsdfEval eval_coated_diffuse(const MatCtx& C, const HkMaterial& m, float3 wo_w, float3 wi_w, float3 n, float4 lam) {
    CoatParams P = coat_params(m, false);
    Spec refl = mat_spec(C, m, 0, lam);
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
        IfaceSample wis = exit_bottom ? base_sample(wi, make_float2(b1, b2), refl, HK_TRANS)
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
                        if (exit_bottom) { fe = base_eval(-w, wi, refl); ep = base_pdf(-w, wi); }
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
                IfaceSample bs = exit_bottom ? base_sample(-w, make_float2(c1, c2), refl, HK_REFL)
                                             : coat_sample(-w, c0, make_float2(c1, c2), ax, ay, eta, HK_REFL);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * fabsf(bs.wi.z) / bs.pdf;
                w = bs.wi;
            } else {
                const bool at_top = z == th;
                const bool ne_spec = at_top ? smooth : false;
                if (!ne_spec) {
                    Spec fn = at_top ? coat_eval(-w, -wis.wi, ax, ay, eta) : base_eval(-w, -wis.wi, refl);
                    if (sp_maxc(fn) > 0.0f) {
                        float wt = 1.0f;
                        if (!exit_bottom || !smooth) {
                            float np = at_top ? coat_pdf(-w, -wis.wi, ax, ay, eta, HK_RT_ALL) : base_pdf(-w, -wis.wi);
                            wt = power_heur(wis.pdf, np);
                        }
                        acc = acc + beta * fn * fabsf(wis.wi.z) * wt * layer_tr(th, wis.wi) * wis.f / wis.pdf;
                    }
                }
                float c0 = pcg32_f32(rng), c1 = pcg32_f32(rng), c2 = pcg32_f32(rng);
                IfaceSample bs = at_top ? coat_sample(-w, c0, make_float2(c1, c2), ax, ay, eta, HK_REFL)
                                        : base_sample(-w, make_float2(c1, c2), refl, HK_REFL);
                if (!bs.valid || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                beta = beta * bs.f * fabsf(bs.wi.z) / bs.pdf;
                w = bs.wi;
                if (!smooth || exit_bottom) {
                    Spec f3e = exit_bottom ? base_eval(-w, wi, refl) : coat_eval(-w, wi, ax, ay, eta);
                    if (sp_maxc(f3e) > 0.0f) {
                        float wt3 = 1.0f;
                        if (!ne_spec) wt3 = power_heur(bs.pdf, exit_bottom ? base_pdf(-w, wi) : coat_pdf(-w, wi, ax, ay, eta, HK_TRANS));
                        acc = acc + beta * layer_tr(th, bs.wi) * f3e * wt3;
                    }
                }
            }
        }
    }
    acc = acc / (float)P.n_samples;
    return eval_make(acc, coated_pdf(wo, wi, P, refl));
}
