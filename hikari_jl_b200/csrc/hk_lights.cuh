// hk_lights.cuh — spectral light sampling, equal-area environment map, BVH light sampler (device side).
// Reference: src/integrators/physical-wavefront/lights.jl:39-290,408-467, src/textures/environment_map.jl:78-160,
// 290-371, src/sampler/sampling.jl:270-360, src/lights/bvh-light-sampler.jl:58-232, src/lights/light-bounds.jl:96-109,166-171,
// src/lights/diffuse-area.jl:54-66.
#pragma once
#include "hk_spectral.cuh"

struct DevEnvMap {
    const float* __restrict__ rgb; int32_t w, h; float rot[9]; float scale_rgb[3];
    const float* __restrict__ cfunc; const float* __restrict__ ccdf; const float* __restrict__ cfint;
    const float* __restrict__ mfunc; const float* __restrict__ mcdf; float mfint; int32_t nu, nv;
};
struct LightCtx {
    DevTables T;
    const HkLight* __restrict__ lights; int32_t n_lights;
    const DevEnvMap* __restrict__ envmaps;
    const struct DevLNode* __restrict__ nodes; const uint32_t* __restrict__ bit_trails; const int32_t* __restrict__ inf_idx;
    int32_t n_infinite, n_bvh;
    // ascending indices of the lights an escaped ray can see (Environment, Ambient): the reference walks the whole light
    // tuple per escaped ray (lights.jl:408-467), which with 10 000 area lights (C3) is 10 002 type tests per ray
    const int32_t* __restrict__ esc_idx; int32_t n_esc;
};
struct LightSample { Spec Li; float3 wi; float pdf; float3 p_light; bool delta; };
HK_DEV LightSample ls_none() { LightSample s; s.Li = sp(0.0f); s.wi = f3(0, 0, 1); s.pdf = 0.0f; s.p_light = f3(0, 0, 0); s.delta = false; return s; }

HK_DEV float2 sphere_to_square(float3 d) {   // environment_map.jl:78-123
    float x = fabsf(d.x), y = fabsf(d.y), z = fabsf(d.z);
    float r = sqrtf(1.0f - z);
    float a = fmaxf(x, y);
    float b = a == 0.0f ? 0.0f : fminf(x, y) / a;
    float phi = 0.406758566246788489601959989e-5f + b * (0.636226545274016134946890922156f + b * (0.61572017898280213493197203466e-2f + b * (-0.247333733281268944196501420480f + b * (0.881770664775316294736387951347e-1f + b * (0.419038818029165735901852432784e-1f + b * -0.251390972343483509333252996350e-1f)))));
    if (x < y) phi = 1.0f - phi;
    float v = phi * r, u = r - v;
    if (d.z < 0.0f) { float t = u; u = v; v = t; u = 1.0f - u; v = 1.0f - v; }
    u = copysignf(u, d.x); v = copysignf(v, d.y);
    return make_float2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
}
HK_DEV float3 square_to_sphere(float2 p) {   // :133-160
    float u = 2.0f * p.x - 1.0f, v = 2.0f * p.y - 1.0f;
    float up = fabsf(u), vp = fabsf(v);
    float sd = 1.0f - (up + vp);
    float r = 1.0f - fabsf(sd);
    float phi = (r == 0.0f ? 1.0f : (vp - up) / r + 1.0f) * HK_PI / 4.0f;
    float z = copysignf(1.0f - r * r, sd);
    float cp = copysignf(dm_cosf(phi), u), sn = copysignf(dm_sinf(phi), v);
    float rc = r * sqrtf(2.0f - r * r);
    return f3(cp * rc, sn * rc, z);
}
HK_DEV float3 mat3_mul(const float* m, float3 v) { return f3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z); }
HK_DEV float3 mat3_tmul(const float* m, float3 v) { return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }
HK_DEV float3 env_texel(const DevEnvMap& E, int y1, int x1) {
    const float* p = E.rgb + ((size_t)(y1 - 1) * E.w + (x1 - 1)) * 3;
    return f3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
HK_DEV float3 env_lookup_dir(const DevEnvMap& E, float3 dir) {   // :290-334
    float2 uv = sphere_to_square(mat3_tmul(E.rot, dir));
    float x = uv.x * (float)(E.w - 1) + 1.0f, y = uv.y * (float)(E.h - 1) + 1.0f;
    int fx0 = floor_i(x), fy0 = floor_i(y);
    int x0 = clampi(fx0, 1, E.w), x1 = clampi(fx0 + 1, 1, E.w), y0 = clampi(fy0, 1, E.h), y1 = clampi(fy0 + 1, 1, E.h);
    x1 = x1 > E.w ? 1 : x1;
    float fx = x - (float)fx0, fy = y - (float)fy0;
    float3 c00 = env_texel(E, y0, x0), c10 = env_texel(E, y0, x1), c01 = env_texel(E, y1, x0), c11 = env_texel(E, y1, x1);
    float3 c0 = c00 * (1.0f - fx) + c10 * fx, c1 = c01 * (1.0f - fx) + c11 * fx;
    return c0 * (1.0f - fy) + c1 * fy;
}
HK_DEV float3 env_lookup_uv(const DevEnvMap& E, float2 uv) {   // :358-371
    return env_texel(E, clampi(floor_i(uv.y * (float)E.h) + 1, 1, E.h), clampi(floor_i(uv.x * (float)E.w) + 1, 1, E.w));
}
HK_DEV int cdf_interval(const float* __restrict__ cdf, int n, float u) {   // sampling.jl:316-345
    int lo = 1, hi = n;
#pragma unroll 1
    for (int it = 0; it < 20; it++) {
        int mid = (lo + hi + 1) / 2;
        bool c = __ldg(cdf + mid - 1) <= u;
        lo = c ? mid : lo; hi = c ? hi : mid - 1;
    }
    return lo;
}
HK_NI_LIGHTS float2 env_sample_uv(const DevEnvMap& E, float2 u, float& pdf) {   // sampling.jl:270-311
    int vo = clampi(cdf_interval(E.mcdf, E.nv + 1, u.y), 1, E.nv);
    float m0 = __ldg(E.mcdf + vo - 1), m1 = __ldg(E.mcdf + vo);
    float dv = u.y - m0, dn = m1 - m0;
    if (dn > 0.0f) dv /= dn;
    float vs = ((float)(vo - 1) + dv) / (float)E.nv;
    float pv = E.mfint > 0.0f ? __ldg(E.mfunc + vo - 1) / E.mfint : 0.0f;
    const float* cc = E.ccdf + (size_t)(vo - 1) * (E.nu + 1);
    int uo = clampi(cdf_interval(cc, E.nu + 1, u.x), 1, E.nu);
    float c0 = __ldg(cc + uo - 1), c1 = __ldg(cc + uo);
    float du = u.x - c0, dnu = c1 - c0;
    if (dnu > 0.0f) du /= dnu;
    float us = ((float)(uo - 1) + du) / (float)E.nu;
    float fi = __ldg(E.cfint + vo - 1);
    float pu = fi > 0.0f ? __ldg(E.cfunc + (size_t)(vo - 1) * E.nu + (uo - 1)) / fi : 0.0f;
    pdf = pu * pv;
    return make_float2(us, vs);
}
HK_DEV float env_pdf_uv(const DevEnvMap& E, float2 uv) {   // sampling.jl:351-360
    int iu = clampi(floor_i(uv.x * (float)E.nu) + 1, 1, E.nu), iv = clampi(floor_i(uv.y * (float)E.nv) + 1, 1, E.nv);
    return __ldg(E.cfunc + (size_t)(iv - 1) * E.nu + (iu - 1)) / E.mfint;
}

HK_DEV Spec light_spectrum(const DevTables& T, const HkLight& L, float4 lam) {
    if (L.spectrum_kind == HK_SPECTRUM_ILLUMINANT) return illuminant_eval(T, Poly3{L.poly[0], L.poly[1], L.poly[2]}, L.illum_scale, lam);
    const float4 q = T.light_pre ? __ldg(T.light_pre + 2 * (&L - (const HkLight*)T.light_base)) : make_pre_illuminant(T, L.rgb[0], L.rgb[1], L.rgb[2]);
    return pre_illuminant(T, q, lam);
}
HK_DEV Spec arealight_Le(const DevTables& T, const HkLight& L, float3 wo, float3 n, float4 lam) {   // diffuse-area.jl:54-66
    if (L.type != HK_LIGHT_DIFFUSE_AREA) return sp(0.0f);
    if (!L.two_sided && dot3(wo, n) < 0.0f) return sp(0.0f);
    const float4 q = T.light_pre ? __ldg(T.light_pre + 2 * (&L - (const HkLight*)T.light_base) + 1) : make_pre_bounded(T, L.rgb[0] * L.scale, L.rgb[1] * L.scale, L.rgb[2] * L.scale);
    return pre_bounded(q, lam);
}
HK_NI_LIGHTS LightSample sample_light(const LightCtx& C, const HkLight& L, float3 p, float4 lam, float2 u) {   // lights.jl:39-290
    LightSample s;
    switch (L.type) {
        case HK_LIGHT_POINT: case HK_LIGHT_SPOT: {
            float3 pos = f3(L.position[0], L.position[1], L.position[2]);
            float3 tl = pos - p;
            float d2 = dot3(tl, tl), d = sqrtf(d2);
            if (d < 1.0e-6f) return ls_none();
            float3 wi = tl / d;
            Spec I = L.scale * light_spectrum(C.T, L, lam);
            if (L.type == HK_LIGHT_SPOT) {
                float ct = norm3(xf_vec(L.world_to_light, -wi)).z;
                if (ct < L.cos_total_width) return ls_none();
                float fall = 1.0f;
                if (ct < L.cos_falloff_start) { float dl = (ct - L.cos_total_width) / (L.cos_falloff_start - L.cos_total_width); fall = dl * dl * dl * dl; }
                I = I * fall;
            }
            s.Li = I / d2; s.wi = wi; s.pdf = 1.0f; s.p_light = pos; s.delta = true;
            return s;
        }
        case HK_LIGHT_DIRECTIONAL: case HK_LIGHT_SUN: {
            float3 wi = -f3(L.direction[0], L.direction[1], L.direction[2]);
            s.Li = L.scale * light_spectrum(C.T, L, lam); s.wi = wi; s.pdf = 1.0f; s.p_light = p + 1.0e6f * wi; s.delta = true;
            return s;
        }
        case HK_LIGHT_ENVIRONMENT: {
            const DevEnvMap& E = C.envmaps[L.env_map - 1];
            float mp;
            float2 uv = env_sample_uv(E, u, mp);
            float3 wi = mat3_mul(E.rot, square_to_sphere(uv));
            float pdf = mp / (4.0f * HK_PI);
            if (pdf <= 0.0f) return ls_none();
            float3 c = env_lookup_uv(E, uv);
            s.Li = uplift_rgb_illuminant(C.T, c.x * E.scale_rgb[0], c.y * E.scale_rgb[1], c.z * E.scale_rgb[2], lam);
            s.wi = wi; s.pdf = pdf; s.p_light = p + 1.0e6f * wi; s.delta = false;
            return s;
        }
        case HK_LIGHT_AMBIENT: {
            float z = 1.0f - 2.0f * u.x;
            float r = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = 2.0f * HK_PI * u.y;
            float3 wi = f3(r * dm_cosf(phi), r * dm_sinf(phi), z);
            s.Li = L.scale * light_spectrum(C.T, L, lam); s.wi = wi; s.pdf = 1.0f / (4.0f * HK_PI); s.p_light = p + 1.0e6f * wi; s.delta = false;
            return s;
        }
        case HK_LIGHT_DIFFUSE_AREA: {
            float b0, b1;
            if (u.x < u.y) { b0 = u.x / 2.0f; b1 = u.y - b0; } else { b1 = u.y / 2.0f; b0 = u.x - b1; }
            float b2 = 1.0f - b0 - b1;
            float3 pl = b0 * f3(L.v[0], L.v[1], L.v[2]) + b1 * f3(L.v[3], L.v[4], L.v[5]) + b2 * f3(L.v[6], L.v[7], L.v[8]);
            float3 tl = pl - p;
            float d2 = dot3(tl, tl);
            if (d2 < 1.0e-12f) return ls_none();
            float d = sqrtf(d2);
            float3 wi = tl / d;
            float3 nl = f3(L.normal[0], L.normal[1], L.normal[2]);
            float ct = fabsf(dot3(nl, -wi));
            if (ct < 1.0e-6f) return ls_none();
            Spec Le = arealight_Le(C.T, L, -wi, nl, lam);
            if (sp_black(Le)) return ls_none();
            s.Li = Le; s.wi = wi; s.pdf = d2 / (ct * L.area); s.p_light = pl; s.delta = false;
            return s;
        }
    }
    return ls_none();
}
HK_NI_LIGHTS Spec escaped_Le(const LightCtx& C, float3 d, float4 lam) {   // lights.jl:408-448
    Spec sum = sp(0.0f);
    for (int k = 0; k < C.n_esc; k++) {
        const HkLight& L = C.lights[__ldg(C.esc_idx + k)];
        if (L.type == HK_LIGHT_ENVIRONMENT) {
            const DevEnvMap& E = C.envmaps[L.env_map - 1];
            float3 c = env_lookup_dir(E, d);
            sum = sum + uplift_rgb_illuminant(C.T, c.x * E.scale_rgb[0], c.y * E.scale_rgb[1], c.z * E.scale_rgb[2], lam);
        } else if (L.type == HK_LIGHT_AMBIENT) sum = sum + L.scale * light_spectrum(C.T, L, lam);
    }
    return sum;
}
HK_NI_LIGHTS float env_light_pdf(const LightCtx& C, float3 d) {   // lights.jl:452-467
    float sum = 0.0f;
    for (int k = 0; k < C.n_esc; k++) {
        const HkLight& L = C.lights[__ldg(C.esc_idx + k)];
        if (L.type == HK_LIGHT_ENVIRONMENT) {
            const DevEnvMap& E = C.envmaps[L.env_map - 1];
            sum = sum + env_pdf_uv(E, sphere_to_square(mat3_tmul(E.rot, d))) / (4.0f * HK_PI);
        }
    }
    return sum;
}

// ---- BVH light sampler ----------------------------------------------------------------------------------
HK_DEV float cos_sub_clamped(float sa, float ca, float sb, float cb) { return ca > cb ? 1.0f : ca * cb + sa * sb; }
HK_DEV float sin_sub_clamped(float sa, float ca, float sb, float cb) { return ca > cb ? 0.0f : sa * cb - ca * sb; }
// Device form of a light-BVH node (64 bytes, four 16-byte loads): everything node_importance needs that does NOT depend on the
// shading point is evaluated once per upload (k_prepare_lnodes, the reference's own operations in the reference's order, so the
// importance keeps its bits): the box centre, half its diagonal, the squared bounding-sphere radius and sin(theta_o).  The box
// corners themselves are not needed after that.  [pc.xyz, half_diag] [w.xyz, phi] [cos_o, sin_o, cos_e, r2] [two_sided, child, leaf, -]
struct DevLNode { float4 a, b, c, d; };
struct LNode { float3 pc, w; float half_diag, phi, cos_o, sin_o, cos_e, r2; uint32_t two_sided, child, leaf; };
HK_DEV DevLNode prepare_lnode(const HkLightBVHNode& N) {
    const float3 lo = f3(N.bounds_min[0], N.bounds_min[1], N.bounds_min[2]), hi = f3(N.bounds_max[0], N.bounds_max[1], N.bounds_max[2]);
    const float3 pc = (lo + hi) * 0.5f;
    const float half_diag = len3(hi - lo) * 0.5f;
    const float3 dr = hi - pc;
    const float r2 = dot3(dr, dr);
    const float so = sqrtf(fmaxf(0.0f, 1.0f - N.cos_theta_o * N.cos_theta_o));
    DevLNode d;
    d.a = make_float4(pc.x, pc.y, pc.z, half_diag); d.b = make_float4(N.w[0], N.w[1], N.w[2], N.phi);
    d.c = make_float4(N.cos_theta_o, so, N.cos_theta_e, r2);
    d.d = make_float4(__uint_as_float(N.two_sided), __uint_as_float(N.child1_or_light_idx), __uint_as_float(N.is_leaf), 0.0f);
    return d;
}
HK_DEV LNode load_lnode(const DevLNode* __restrict__ nodes, int idx1) {
    const float4* q = reinterpret_cast<const float4*>(nodes + (idx1 - 1));
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    LNode n;
    n.pc = f3(a.x, a.y, a.z); n.half_diag = a.w; n.w = f3(b.x, b.y, b.z); n.phi = b.w;
    n.cos_o = c.x; n.sin_o = c.y; n.cos_e = c.z; n.r2 = c.w;
    n.two_sided = __float_as_uint(d.x); n.child = __float_as_uint(d.y); n.leaf = __float_as_uint(d.z);
    return n;
}
HK_NI_LIGHTS float lnode_importance(const LNode& N, float3 p, float3 n) {   // bvh-light-sampler.jl:58-91
    if (N.phi == 0.0f) return 0.0f;
    float3 dp = p - N.pc;
    const float d2raw = dot3(dp, dp);
    float d2 = fmaxf(d2raw, N.half_diag);
    float3 wi = norm3(dp);
    float cw = dot3(N.w, wi);
    if (N.two_sided) cw = fabsf(cw);
    float sw = sqrtf(fmaxf(0.0f, 1.0f - cw * cw));
    // bound_subtended_directions (light-bounds.jl:96-109): only cosθ is needed
    float cb = d2raw < N.r2 ? -1.0f : sqrtf(fmaxf(0.0f, 1.0f - N.r2 / d2raw));
    float sb = sqrtf(fmaxf(0.0f, 1.0f - cb * cb));
    float cx = cos_sub_clamped(sw, cw, N.sin_o, N.cos_o), sx = sin_sub_clamped(sw, cw, N.sin_o, N.cos_o);
    float cp = cos_sub_clamped(sx, cx, sb, cb);
    if (cp <= N.cos_e) return 0.0f;
    float imp = N.phi * cp / d2;
    if (n.x != 0.0f || n.y != 0.0f || n.z != 0.0f) {
        float ci = fabsf(dot3(wi, n));
        float si = sqrtf(fmaxf(0.0f, 1.0f - ci * ci));
        imp *= cos_sub_clamped(si, ci, sb, cb);
    }
    return fmaxf(imp, 0.0f);
}
HK_NI_LIGHTS int bvh_sample_light(const LightCtx& C, float3 p, float3 n, float u, float& pmf_out) {   // :105-170
    pmf_out = 0.0f;
    if (C.n_infinite + C.n_bvh == 0) return 0;
    const bool has_bvh = C.n_bvh > 0;
    float p_inf = (float)C.n_infinite / (float)(C.n_infinite + (has_bvh ? 1 : 0));
    if (C.n_infinite > 0 && u < p_inf) {
        int idx = min(floor_i((u / p_inf) * (float)C.n_infinite), C.n_infinite - 1) + 1;
        pmf_out = p_inf / (float)C.n_infinite;
        return __ldg(C.inf_idx + idx - 1);
    }
    if (!has_bvh) return 0;
    float ub = C.n_infinite > 0 ? fminf((u - p_inf) / (1.0f - p_inf), 0.99999994f) : fminf(u, 0.99999994f);
    float pmf = 1.0f - p_inf;
    int ni = 1;
    // (the node being expanded was loaded as a child one level up: its leaf flag / child index are carried over, not re-fetched)
    uint32_t cur_leaf, cur_child;
    { const LNode R = load_lnode(C.nodes, 1); cur_leaf = R.leaf; cur_child = R.child; }
    for (int it = 0; it < 64; it++) {
        if (cur_leaf) { pmf_out = pmf; return (int)cur_child; }
        int c0i = ni + 1, c1i = (int)cur_child;
        const LNode N0 = load_lnode(C.nodes, c0i), N1 = load_lnode(C.nodes, c1i);
        float c0 = lnode_importance(N0, p, n), c1 = lnode_importance(N1, p, n);
        if (c0 == 0.0f && c1 == 0.0f) return 0;
        float p0 = c0 / (c0 + c1);
        if (ub < p0) { pmf *= p0; ub = ub / p0; ni = c0i; cur_leaf = N0.leaf; cur_child = N0.child; }
        else { pmf *= (1.0f - p0); ub = (ub - p0) / (1.0f - p0); ni = c1i; cur_leaf = N1.leaf; cur_child = N1.child; }
    }
    return 0;
}
// Same selection, computed cooperatively by the lanes that arrive together (__activemask): with infinite lights in the
// scene only a fraction of a warp's lanes descends the light BVH (C3: 1/3 choose the 10 000-emitter tree, ncu: 6.8 of 32
// lanes active in the descent), so the descents are served by PAIRS of lanes -- one child importance each, exchanged by
// shuffle -- up to 16 descents per round.  Per item the arithmetic and its order are those of bvh_sample_light: same bits.
// the part of bvh_sample_light ahead of the tree (:105-125): an infinite light is picked uniformly (returns its index, pmf_out set), or
// the light BVH must be descended (need = true; ub = the remapped sample, pmf = the probability of having chosen the tree)
HK_DEV int light_select_prologue(const LightCtx& C, float u, float& pmf_out, bool& need, float& ub, float& pmf) {
    pmf_out = 0.0f; need = false; ub = 0.0f; pmf = 0.0f;
    int result = 0;
    if (C.n_infinite + C.n_bvh != 0) {
        const bool has_bvh = C.n_bvh > 0;
        const float p_inf = (float)C.n_infinite / (float)(C.n_infinite + (has_bvh ? 1 : 0));
        if (C.n_infinite > 0 && u < p_inf) {
            int idx = min(floor_i((u / p_inf) * (float)C.n_infinite), C.n_infinite - 1) + 1;
            pmf_out = p_inf / (float)C.n_infinite;
            result = __ldg(C.inf_idx + idx - 1);
        } else if (has_bvh) {
            need = true;
            ub = C.n_infinite > 0 ? fminf((u - p_inf) / (1.0f - p_inf), 0.99999994f) : fminf(u, 0.99999994f);
            pmf = 1.0f - p_inf;
        }
    }
    return result;
}
HK_NI_LIGHTS int bvh_descend_coop(const LightCtx& C, float3 p, float3 n, bool need, float ub, float pmf, int result, float& pmf_out);
HK_NI_LIGHTS int bvh_sample_light_coop(const LightCtx& C, float3 p, float3 n, float u, float& pmf_out) {
    bool need; float ub, pmf;
    const int result = light_select_prologue(C, u, pmf_out, need, ub, pmf);
    return bvh_descend_coop(C, p, n, need, ub, pmf, result, pmf_out);
}
// the descent for the lanes that arrive together with need = true (the others pass their `result` / pmf_out through)
HK_NI_LIGHTS int bvh_descend_coop(const LightCtx& C, float3 p, float3 n, bool need, float ub, float pmf, int result, float& pmf_out) {
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    unsigned todo = __ballot_sync(m, need);
    if (todo == 0u) return result;
    const unsigned n_helpers = (unsigned)__popc(m), hidx = (unsigned)__popc(m & lt);
    const unsigned n_pairs = n_helpers >> 1, pair = hidx >> 1, child = hidx & 1u;
    if (n_pairs == 0u) {        // a lone lane: plain descent
        if (need) {
            int ni = 1;
            uint32_t cur_leaf, cur_child;
            { const LNode R = load_lnode(C.nodes, 1); cur_leaf = R.leaf; cur_child = R.child; }
            for (int it = 0; it < 64; it++) {
                if (cur_leaf) { pmf_out = pmf; return (int)cur_child; }
                int c0i = ni + 1, c1i = (int)cur_child;
                const LNode N0 = load_lnode(C.nodes, c0i), N1 = load_lnode(C.nodes, c1i);
                float c0 = lnode_importance(N0, p, n), c1 = lnode_importance(N1, p, n);
                if (c0 == 0.0f && c1 == 0.0f) return 0;
                float p0 = c0 / (c0 + c1);
                if (ub < p0) { pmf *= p0; ub = ub / p0; ni = c0i; cur_leaf = N0.leaf; cur_child = N0.child; }
                else { pmf *= (1.0f - p0); ub = (ub - p0) / (1.0f - p0); ni = c1i; cur_leaf = N1.leaf; cur_child = N1.child; }
            }
        }
        return 0;
    }
    const bool paired = pair < n_pairs;                                            // (the last lane of an odd group has no partner)
    const unsigned partner = paired ? __fns(m, 0, (int)(hidx ^ 1u) + 1) : lane;
    while (todo != 0u) {
        const unsigned owner = (paired && pair < (unsigned)__popc(todo)) ? __fns(todo, 0, (int)pair + 1) : 0xFFFFFFFFu;
        const bool active = owner != 0xFFFFFFFFu;
        const unsigned src = active ? owner : lane;
        const float3 bp = f3(__shfl_sync(m, p.x, src), __shfl_sync(m, p.y, src), __shfl_sync(m, p.z, src));
        const float3 bn = f3(__shfl_sync(m, n.x, src), __shfl_sync(m, n.y, src), __shfl_sync(m, n.z, src));
        float bub = __shfl_sync(m, ub, src), bpmf = __shfl_sync(m, pmf, src);
        int ni = 1, r_light = 0;
        float r_pmf = 0.0f;
        bool done = !active;
        // the node being expanded was fetched one level up by one lane of the pair (as the child whose importance it evaluated): its
        // leaf flag and child index travel by shuffle, so a level costs each lane ONE dependent node fetch, not two
        uint32_t cur_leaf = 0u, cur_child = 0u;
        if (active) { const LNode R = load_lnode(C.nodes, 1); cur_leaf = R.leaf; cur_child = R.child; }
        for (int it = 0; it < 64; it++) {
            if (__all_sync(m, done)) break;
            float mine = 0.0f;
            int c1i = 0;
            uint32_t my_leaf = 0u, my_child = 0u;
            if (!done) {
                if (cur_leaf) { r_light = (int)cur_child; r_pmf = bpmf; done = true; }
                else {
                    c1i = (int)cur_child;
                    const LNode Nc = load_lnode(C.nodes, child == 0u ? ni + 1 : c1i);
                    mine = lnode_importance(Nc, bp, bn); my_leaf = Nc.leaf; my_child = Nc.child;
                }
            }
            const float other = __shfl_sync(m, mine, partner);
            const uint32_t o_leaf = __shfl_sync(m, my_leaf, partner), o_child = __shfl_sync(m, my_child, partner);
            if (!done) {
                const float c0 = child == 0u ? mine : other, c1 = child == 0u ? other : mine;
                if (c0 == 0.0f && c1 == 0.0f) done = true;                       // result 0, pmf 0
                else {
                    const float p0 = c0 / (c0 + c1);
                    const bool first = bub < p0;
                    if (first) { bpmf *= p0; bub = bub / p0; ni = ni + 1; }
                    else { bpmf *= (1.0f - p0); bub = (bub - p0) / (1.0f - p0); ni = c1i; }
                    const bool i_hold_it = first == (child == 0u);                // the chosen child is the one this lane fetched
                    cur_leaf = i_hold_it ? my_leaf : o_leaf; cur_child = i_hold_it ? my_child : o_child;
                }
            }
        }
        // hand the results back: the r-th pending owner was served by pair r (its even lane holds the result)
        const unsigned myrank = (unsigned)__popc(todo & lt);
        const bool served = ((todo >> lane) & 1u) != 0u && myrank < n_pairs;
        const unsigned from = served ? __fns(m, 0, (int)(2u * myrank) + 1) : lane;
        const int g_light = __shfl_sync(m, r_light, from);
        const float g_pmf = __shfl_sync(m, r_pmf, from);
        if (served) { result = g_light; pmf_out = g_pmf; }
        todo &= ~__ballot_sync(m, served);
    }
    return result;
}
// small light sets: the serial descent is a level or two and the pairing overhead (8 broadcasts + result shuffles) would
// dominate (C2, 3 lights: shading +45 % with the cooperative form); the switch is uniform per launch
#ifndef HK_COOP_MIN_LIGHTS
#define HK_COOP_MIN_LIGHTS 64
#endif
HK_DEV int bvh_sample_light_auto(const LightCtx& C, float3 p, float3 n, float u, float& pmf_out) {
    return C.n_bvh >= HK_COOP_MIN_LIGHTS ? bvh_sample_light_coop(C, p, n, u, pmf_out) : bvh_sample_light(C, p, n, u, pmf_out);
}
HK_NI_LIGHTS float bvh_light_pmf(const LightCtx& C, float3 p, float3 n, int flat_idx) {   // :184-232
    if (flat_idx < 1) return 0.0f;
    const bool has_bvh = C.n_bvh > 0;
    uint32_t trail = __ldg(C.bit_trails + flat_idx - 1);
    if (trail == 0xFFFFFFFFu) return C.n_infinite == 0 ? 0.0f : 1.0f / (float)(C.n_infinite + (has_bvh ? 1 : 0));
    if (!has_bvh) return 0.0f;
    float pmf = 1.0f - (float)C.n_infinite / (float)(C.n_infinite + 1);
    int ni = 1;
    uint32_t cur_leaf, cur_child;
    { const LNode R = load_lnode(C.nodes, 1); cur_leaf = R.leaf; cur_child = R.child; }
    for (int it = 0; it < 64; it++) {
        if (cur_leaf) return pmf;
        int c0i = ni + 1, c1i = (int)cur_child;
        const LNode N0 = load_lnode(C.nodes, c0i), N1 = load_lnode(C.nodes, c1i);
        float c0 = lnode_importance(N0, p, n), c1 = lnode_importance(N1, p, n);
        float sc = c0 + c1;
        if (sc <= 0.0f) return 0.0f;
        if ((trail & 1u) == 0) { pmf *= c0 / sc; ni = c0i; cur_leaf = N0.leaf; cur_child = N0.child; } else { pmf *= c1 / sc; ni = c1i; cur_leaf = N1.leaf; cur_child = N1.child; }
        trail >>= 1;
    }
    return pmf;
}
