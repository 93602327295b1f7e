// ok_lights.h — ORACLE (test infrastructure, NOT product code).
// Spectral light sampling, environment map, BVH light sampler (sample + pmf).
#pragma once
#include "ok_spectral.h"

namespace ok {

struct LightSample {   // PWLightSample, src/integrators/physical-wavefront/lights.jl:13-27
    Spec Li; V3 wi; float pdf; V3 p_light; bool is_delta;
    LightSample() : Li(0.0f), wi(0, 0, 1), pdf(0.0f), p_light(0, 0, 0), is_delta(false) {}
    LightSample(Spec L, V3 w, float p, V3 pl, bool d) : Li(L), wi(w), pdf(p), p_light(pl), is_delta(d) {}
};

struct LightCtx {
    const Tables* T;
    const HkLight* lights; uint32_t n_lights;
    const HkEnvMap* envmaps; uint32_t n_envmaps;
    const HkLightSampler* sampler;
};

// ---- equal-area mapping, src/textures/environment_map.jl:78-160 --------------------------------
inline V2 equal_area_sphere_to_square(V3 d) {
    float x = std::fabs(d.x), y = std::fabs(d.y), z = std::fabs(d.z);
    float r = std::sqrt(1.0f - z);
    float a = std::max(x, y);
    float b = a == 0.0f ? 0.0f : std::min(x, y) / a;
    const float t1 = 0.406758566246788489601959989e-5f, t2 = 0.636226545274016134946890922156f,
                t3 = 0.61572017898280213493197203466e-2f, t4 = -0.247333733281268944196501420480f,
                t5 = 0.881770664775316294736387951347e-1f, t6 = 0.419038818029165735901852432784e-1f,
                t7 = -0.251390972343483509333252996350e-1f;
    float phi = t1 + b * (t2 + b * (t3 + b * (t4 + b * (t5 + b * (t6 + b * t7)))));
    if (x < y) phi = 1.0f - phi;
    float v = phi * r;
    float u = r - v;
    if (d.z < 0.0f) { float tmp = u; u = v; v = tmp; u = 1.0f - u; v = 1.0f - v; }
    u = std::copysign(u, d.x);
    v = std::copysign(v, d.y);
    return V2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
}
inline V3 equal_area_square_to_sphere(V2 p) {
    float u = 2.0f * p.x - 1.0f, v = 2.0f * p.y - 1.0f;
    float up = std::fabs(u), vp = std::fabs(v);
    float sd = 1.0f - (up + vp);
    float d = std::fabs(sd);
    float r = 1.0f - d;
    float phi = (r == 0.0f ? 1.0f : (vp - up) / r + 1.0f) * PI_F / 4.0f;
    float z = std::copysign(1.0f - r * r, sd);
    float cp = std::copysign(dm_cosf(phi), u);
    float sp = std::copysign(dm_sinf(phi), v);
    float rc = r * std::sqrt(2.0f - r * r);
    return V3(cp * rc, sp * rc, z);
}
// rotation is a column-major Mat3f: M*v and transpose(M)*v  (environment_map.jl:202-218)
inline V3 mat3_mul(const float* m, V3 v) {
    return V3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z);
}
inline V3 mat3_tmul(const float* m, V3 v) {
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
inline void env_texel(const HkEnvMap& E, int y1, int x1, float* out) {   // 1-based [v, u]
    const float* p = E.rgb + ((size_t)(y1 - 1) * E.w + (x1 - 1)) * 3;
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
// environment_map.jl:290-334  bilinear lookup by direction
inline void env_lookup_dir(const HkEnvMap& E, V3 dir, float* out) {
    V2 uv = equal_area_sphere_to_square(mat3_tmul(E.rotation, dir));
    int w = E.w, h = E.h;
    float x = uv.x * (float)(w - 1) + 1.0f;
    float y = uv.y * (float)(h - 1) + 1.0f;
    int x0 = floor_int32(x), y0 = floor_int32(y);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = clampi(x0, 1, w); x1 = clampi(x1, 1, w); y0 = clampi(y0, 1, h); y1 = clampi(y1, 1, h);
    x1 = x1 > w ? 1 : x1;
    float fx = x - (float)floor_int32(x), fy = y - (float)floor_int32(y);
    float c00[3], c10[3], c01[3], c11[3];
    env_texel(E, y0, x0, c00); env_texel(E, y0, x1, c10); env_texel(E, y1, x0, c01); env_texel(E, y1, x1, c11);
    for (int i = 0; i < 3; i++) {
        float c0 = c00[i] * (1.0f - fx) + c10[i] * fx;
        float c1 = c01[i] * (1.0f - fx) + c11[i] * fx;
        out[i] = c0 * (1.0f - fy) + c1 * fy;
    }
}
// environment_map.jl:358-371  nearest lookup by uv
inline void env_lookup_uv(const HkEnvMap& E, V2 uv, float* out) {
    int ui = clampi(floor_int32(uv.x * (float)E.w) + 1, 1, E.w);
    int vi = clampi(floor_int32(uv.y * (float)E.h) + 1, 1, E.h);
    env_texel(E, vi, ui, out);
}
// src/sampler/sampling.jl:316-345  20-step branchless search, 1-based result
inline int find_interval_binary(const float* cdf, int n, float u) {
    int lo = 1, hi = n;
    for (int it = 0; it < 20; it++) {
        int mid = (lo + hi + 1) / 2;
        bool c = cdf[mid - 1] <= u;
        lo = c ? mid : lo;
        hi = c ? hi : mid - 1;
    }
    return lo;
}
// sampling.jl:270-311
inline V2 dist2d_sample_continuous(const HkEnvMap& E, V2 u, float& pdf) {
    int nu = E.nu, nv = E.nv;
    int vo = clampi(find_interval_binary(E.marginal_cdf, nv + 1, u.y), 1, nv);
    float duv = u.y - E.marginal_cdf[vo - 1];
    float dnv = E.marginal_cdf[vo] - E.marginal_cdf[vo - 1];
    if (dnv > 0.0f) duv /= dnv;
    float v_s = ((float)(vo - 1) + duv) / (float)nv;
    float pdf_v = E.marginal_func_int > 0.0f ? E.marginal_func[vo - 1] / E.marginal_func_int : 0.0f;
    const float* ccdf = E.conditional_cdf + (size_t)(vo - 1) * (nu + 1);
    int uo = clampi(find_interval_binary(ccdf, nu + 1, u.x), 1, nu);
    float duu = u.x - ccdf[uo - 1];
    float dnu = ccdf[uo] - ccdf[uo - 1];
    if (dnu > 0.0f) duu /= dnu;
    float u_s = ((float)(uo - 1) + duu) / (float)nu;
    float fiv = E.conditional_func_int[vo - 1];
    float pdf_u = fiv > 0.0f ? E.conditional_func[(size_t)(vo - 1) * nu + (uo - 1)] / fiv : 0.0f;
    pdf = pdf_u * pdf_v;
    return V2(u_s, v_s);
}
// sampling.jl:351-360
inline float dist2d_pdf(const HkEnvMap& E, V2 uv) {
    int iu = clampi(floor_int32(uv.x * (float)E.nu) + 1, 1, E.nu);
    int iv = clampi(floor_int32(uv.y * (float)E.nv) + 1, 1, E.nv);
    return E.conditional_func[(size_t)(iv - 1) * E.nu + (iu - 1)] / E.marginal_func_int;
}

// ---- arealight_Le, src/lights/diffuse-area.jl:54-66 -----------------------------------------------
inline Spec arealight_Le(const LightCtx& C, const HkLight& L, V3 wo, V3 n, const Wavelengths& l) {
    if (L.type != HK_LIGHT_DIFFUSE_AREA) return Spec();
    if (!L.two_sided && dot(wo, n) < 0.0f) return Spec();
    float le[3] = {L.rgb[0] * L.scale, L.rgb[1] * L.scale, L.rgb[2] * L.scale};
    return uplift_rgb(*C.T, le, l);
}

// ---- sample_light_spectral, lights.jl:39-290 ------------------------------------------------------
inline LightSample sample_light(const LightCtx& C, const HkLight& L, V3 p, const Wavelengths& l, V2 u) {
    const Tables& T = *C.T;
    switch (L.type) {
        case HK_LIGHT_POINT: {
            V3 pos(L.position[0], L.position[1], L.position[2]);
            V3 tl = pos - p;
            float d2 = dot(tl, tl);
            float d = std::sqrt(d2);
            if (d < 1.0e-6f) return LightSample();
            V3 wi = tl / d;
            Spec Li = L.scale * sample_light_spectrum(T, L, l) / d2;
            return LightSample(Li, wi, 1.0f, pos, true);
        }
        case HK_LIGHT_SPOT: {
            V3 pos(L.position[0], L.position[1], L.position[2]);
            V3 tl = pos - p;
            float d2 = dot(tl, tl);
            float d = std::sqrt(d2);
            if (d < 1.0e-6f) return LightSample();
            V3 wi = tl / d;
            V3 wl = normalize(xform_vec(L.world_to_light, -wi));
            float ct = wl.z;
            if (ct < L.cos_total_width) return LightSample();
            float fall;
            if (ct >= L.cos_falloff_start) fall = 1.0f;
            else { float dl = (ct - L.cos_total_width) / (L.cos_falloff_start - L.cos_total_width); fall = dl * dl * dl * dl; }
            Spec Li = L.scale * sample_light_spectrum(T, L, l) * fall / d2;
            return LightSample(Li, wi, 1.0f, pos, true);
        }
        case HK_LIGHT_DIRECTIONAL:
        case HK_LIGHT_SUN: {
            V3 wi = -V3(L.direction[0], L.direction[1], L.direction[2]);
            V3 pl = p + 1.0e6f * wi;
            // lights.jl:125,147 call uplift_rgb_illuminant(table, light.i, lambda): RGB -> uplift, baked -> Sample
            Spec Li = L.scale * sample_light_spectrum(T, L, l);
            return LightSample(Li, wi, 1.0f, pl, true);
        }
        case HK_LIGHT_ENVIRONMENT: {
            const HkEnvMap& E = C.envmaps[L.env_map - 1];
            float map_pdf;
            V2 uv = dist2d_sample_continuous(E, u, map_pdf);
            V3 wi = mat3_mul(E.rotation, equal_area_square_to_sphere(uv));
            float pdf = map_pdf / (4.0f * PI_F);
            if (pdf <= 0.0f) return LightSample();
            float rgb[3]; env_lookup_uv(E, uv, rgb);
            for (int i = 0; i < 3; i++) rgb[i] = rgb[i] * E.scale_rgb[i];
            V3 pl = p + 1.0e6f * wi;
            Spec Li = uplift_rgb_illuminant(T, rgb, l);
            return LightSample(Li, wi, pdf, pl, false);
        }
        case HK_LIGHT_AMBIENT: {
            float z = 1.0f - 2.0f * u.x;
            float r = std::sqrt(std::max(0.0f, 1.0f - z * z));
            float phi = 2.0f * PI_F * u.y;
            V3 wi(r * dm_cosf(phi), r * dm_sinf(phi), z);
            float pdf = 1.0f / (4.0f * PI_F);
            V3 pl = p + 1.0e6f * wi;
            Spec Li = L.scale * sample_light_spectrum(T, L, l);
            return LightSample(Li, wi, pdf, pl, false);
        }
        case HK_LIGHT_DIFFUSE_AREA: {
            float b0, b1;
            if (u.x < u.y) { b0 = u.x / 2.0f; b1 = u.y - b0; }
            else { b1 = u.y / 2.0f; b0 = u.x - b1; }
            float b2 = 1.0f - b0 - b1;
            V3 v0(L.v[0], L.v[1], L.v[2]), v1(L.v[3], L.v[4], L.v[5]), v2(L.v[6], L.v[7], L.v[8]);
            V3 pl = b0 * v0 + b1 * v1 + b2 * v2;
            V3 tl = pl - p;
            float d2 = dot(tl, tl);
            if (d2 < 1.0e-12f) return LightSample();
            float d = std::sqrt(d2);
            V3 wi = tl / d;
            V3 nl(L.normal[0], L.normal[1], L.normal[2]);
            float ct = std::fabs(dot(nl, -wi));
            if (ct < 1.0e-6f) return LightSample();
            float pdf = d2 / (ct * L.area);
            V3 wo(-wi.x, -wi.y, -wi.z);
            Spec Le = arealight_Le(C, L, wo, nl, l);
            if (is_black(Le)) return LightSample();
            return LightSample(Le, wi, pdf, pl, false);
        }
    }
    return LightSample();
}

// lights.jl:408-448  Σ over environment-type lights
inline Spec evaluate_escaped_ray(const LightCtx& C, V3 d, const Wavelengths& l) {
    Spec sum(0.0f);
    for (uint32_t i = 0; i < C.n_lights; i++) {
        const HkLight& L = C.lights[i];
        if (L.type == HK_LIGHT_ENVIRONMENT) {
            const HkEnvMap& E = C.envmaps[L.env_map - 1];
            float rgb[3]; env_lookup_dir(E, d, rgb);
            for (int k = 0; k < 3; k++) rgb[k] = rgb[k] * E.scale_rgb[k];
            sum = sum + uplift_rgb_illuminant(*C.T, rgb, l);
        } else if (L.type == HK_LIGHT_AMBIENT) {
            sum = sum + L.scale * sample_light_spectrum(*C.T, L, l);
        } else {
            sum = sum + Spec(0.0f);
        }
    }
    return sum;
}
// lights.jl:452-467
inline float compute_env_light_pdf(const LightCtx& C, V3 d) {
    float sum = 0.0f;
    for (uint32_t i = 0; i < C.n_lights; i++) {
        const HkLight& L = C.lights[i];
        if (L.type == HK_LIGHT_ENVIRONMENT) {
            const HkEnvMap& E = C.envmaps[L.env_map - 1];
            V2 uv = equal_area_sphere_to_square(mat3_tmul(E.rotation, d));
            sum = sum + dist2d_pdf(E, uv) / (4.0f * PI_F);
        } else {
            sum = sum + 0.0f;
        }
    }
    return sum;
}

// ---- BVH light sampler, src/lights/bvh-light-sampler.jl:58-232 + light-bounds.jl:96-109,166-171 ---
inline float cos_sub_clamped(float sa, float ca, float sb, float cb) { return ca > cb ? 1.0f : ca * cb + sa * sb; }
inline float sin_sub_clamped(float sa, float ca, float sb, float cb) { return ca > cb ? 0.0f : sa * cb - ca * sb; }
inline float dist2(V3 a, V3 b) { V3 d = a - b; return dot(d, d); }
inline float bound_subtended_cos(V3 bmin, V3 bmax, V3 p) {
    V3 pc = (bmin + bmax) * 0.5f;
    float r2 = dist2(bmax, pc);
    float d2 = dist2(p, pc);
    if (d2 < r2) return -1.0f;
    float s2 = r2 / d2;
    return std::sqrt(std::max(0.0f, 1.0f - s2));
}
inline float node_importance(const HkLightBVHNode& N, V3 p, V3 n) {
    if (N.phi == 0.0f) return 0.0f;
    V3 bmin(N.bounds_min[0], N.bounds_min[1], N.bounds_min[2]), bmax(N.bounds_max[0], N.bounds_max[1], N.bounds_max[2]);
    V3 w(N.w[0], N.w[1], N.w[2]);
    V3 pc = (bmin + bmax) * 0.5f;
    float d2 = dist2(p, pc);
    d2 = std::max(d2, norm(bmax - bmin) * 0.5f);
    V3 wi = normalize(p - pc);
    float cw = dot(w, wi);
    if (N.two_sided) cw = std::fabs(cw);
    float sw = std::sqrt(std::max(0.0f, 1.0f - cw * cw));
    float cb = bound_subtended_cos(bmin, bmax, p);
    float sb = std::sqrt(std::max(0.0f, 1.0f - cb * cb));
    float so = std::sqrt(std::max(0.0f, 1.0f - N.cos_theta_o * N.cos_theta_o));
    float cx = cos_sub_clamped(sw, cw, so, N.cos_theta_o);
    float sx = sin_sub_clamped(sw, cw, so, N.cos_theta_o);
    float cp = cos_sub_clamped(sx, cx, sb, cb);
    if (cp <= N.cos_theta_e) return 0.0f;
    float imp = N.phi * cp / d2;
    if (n != V3(0.0f)) {
        float ci = std::fabs(dot(wi, n));
        float si = std::sqrt(std::max(0.0f, 1.0f - ci * ci));
        imp *= cos_sub_clamped(si, ci, sb, cb);
    }
    return std::max(imp, 0.0f);
}
inline int32_t bvh_sample_light(const HkLightSampler& S, V3 p, V3 n, float u, float& pmf_out) {
    int32_t ninf = (int32_t)S.n_infinite, nbvh = (int32_t)S.n_bvh_lights;
    pmf_out = 0.0f;
    if (ninf + nbvh == 0) return 0;
    bool has_bvh = nbvh > 0;
    float p_inf = (float)ninf / (float)(ninf + (has_bvh ? 1 : 0));
    if (ninf > 0 && u < p_inf) {
        float ur = u / p_inf;
        int32_t idx = std::min(floor_int32(ur * (float)ninf), ninf - 1) + 1;
        pmf_out = p_inf / (float)ninf;
        return S.infinite_light_indices[idx - 1];
    }
    if (!has_bvh) return 0;
    float ub = ninf > 0 ? std::min((u - p_inf) / (1.0f - p_inf), 0.99999994f) : std::min(u, 0.99999994f);
    float pmf = 1.0f - p_inf;
    int32_t ni = 1;
    for (int it = 0; it < 64; it++) {
        const HkLightBVHNode& N = S.nodes[ni - 1];
        if (N.is_leaf) { pmf_out = pmf; return (int32_t)N.child1_or_light_idx; }
        int32_t c0i = ni + 1, c1i = (int32_t)N.child1_or_light_idx;
        float c0 = node_importance(S.nodes[c0i - 1], p, n), c1 = node_importance(S.nodes[c1i - 1], p, n);
        if (c0 == 0.0f && c1 == 0.0f) return 0;
        float p0 = c0 / (c0 + c1);
        if (ub < p0) { pmf *= p0; ub = ub / p0; ni = c0i; }
        else { pmf *= (1.0f - p0); ub = (ub - p0) / (1.0f - p0); ni = c1i; }
    }
    return 0;
}
inline float bvh_pmf(const HkLightSampler& S, V3 p, V3 n, int32_t flat_idx) {
    if (flat_idx < 1) return 0.0f;
    int32_t ninf = (int32_t)S.n_infinite, nbvh = (int32_t)S.n_bvh_lights;
    bool has_bvh = nbvh > 0;
    uint32_t trail = S.light_to_bit_trail[flat_idx - 1];
    if (trail == 0xFFFFFFFFu) {
        if (ninf == 0) return 0.0f;
        return 1.0f / (float)(ninf + (has_bvh ? 1 : 0));
    }
    if (!has_bvh) return 0.0f;
    float p_inf = (float)ninf / (float)(ninf + 1);
    float pmf = 1.0f - p_inf;
    int32_t ni = 1;
    for (int it = 0; it < 64; it++) {
        const HkLightBVHNode& N = S.nodes[ni - 1];
        if (N.is_leaf) return pmf;
        int32_t c0i = ni + 1, c1i = (int32_t)N.child1_or_light_idx;
        float c0 = node_importance(S.nodes[c0i - 1], p, n), c1 = node_importance(S.nodes[c1i - 1], p, n);
        float sc = c0 + c1;
        if (sc <= 0.0f) return 0.0f;
        if ((trail & 1u) == 0) { pmf *= c0 / sc; ni = c0i; }
        else { pmf *= c1 / sc; ni = c1i; }
        trail >>= 1;
    }
    return pmf;
}

}  // namespace ok
