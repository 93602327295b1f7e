// ok_volpath.h — ORACLE (test infrastructure, NOT product code).
// CPU restatement of the VolPath wavefront render loop with the reference's AOS work items and
// stage order (src/integrators/volpath/{volpath,workitems,intersection,surface-eval}.jl).
// Queue pushes happen in item order (deterministic); stages are parallelised with OpenMP over items,
// which is safe because every pixel owns at most one path per sample pass.
#pragma once
#include "ok_core.h"
#include "ok_spectral.h"
#include "ok_bsdf.h"
#include "ok_lights.h"
#include "ok_accel.h"
#include "ok_media.h"
#include <cstring>
#include <vector>
#include <string>

namespace ok {

// ---- work items (workitems.jl) -------------------------------------------------------------------
struct RayWork {          // VPRayWorkItem :38-52
    Ray ray; int32_t depth; Wavelengths lambda; int32_t pixel_index;
    Spec beta, r_u, r_l; V3 prev_p, prev_n; float eta_scale; bool specular_bounce, any_non_specular; uint32_t medium;
};
struct SurfGeom { V3 pi, n, dpdu, dpdv, ns, dpdus, dpdvs; V2 uv; };
struct HitSurfaceWork {   // VPHitSurfaceWorkItem :326-368
    Ray ray; SurfGeom g; uint32_t material; HkMediumInterface iface; uint32_t face_idx; float bary[3];
    uint32_t arealight_flat_idx; float triangle_area;
    Wavelengths lambda; int32_t pixel_index; Spec beta, r_u, r_l; int32_t depth; float eta_scale;
    bool specular_bounce, any_non_specular; V3 prev_p, prev_n; uint32_t current_medium; float t_hit;
};
struct MaterialEvalWork { // VPMaterialEvalWorkItem :209-246
    SurfGeom g; float time; uint32_t face_idx; float bary[3]; V3 wo; uint32_t material; HkMediumInterface iface;
    Wavelengths lambda; int32_t pixel_index; Spec beta, r_u, r_l; int32_t depth; float eta_scale;
    bool specular_bounce, any_non_specular; V3 prev_p, prev_n; uint32_t current_medium;
};
struct ShadowWork {       // VPShadowRayWorkItem :259-268
    Ray ray; float t_max; Wavelengths lambda; Spec Ld, r_u, r_l; int32_t pixel_index; uint32_t medium;
};
struct EscapedWork {      // VPEscapedRayWorkItem :279-290
    V3 ray_d; Wavelengths lambda; int32_t pixel_index; Spec beta, r_u, r_l; int32_t depth; bool specular_bounce; V3 prev_p, prev_n;
};
struct MediumSampleWork { // VPMediumSampleWorkItem :91-129
    RayWork w; float t_max; bool has_surface_hit; SurfGeom g; uint32_t material; HkMediumInterface iface;
    uint32_t face_idx; float bary[3]; uint32_t arealight_flat_idx; float triangle_area;
};
struct MediumScatterWork {// VPMediumScatterWorkItem :182-193
    V3 p, wo; float time; Wavelengths lambda; int32_t pixel_index; Spec beta, r_u; int32_t depth; uint32_t medium; float g;
};

template <class T> struct Opt { bool valid = false; T v; };
template <class T> static void compact(const std::vector<Opt<T>>& in, std::vector<T>& out) {
    for (const auto& o : in) if (o.valid) out.push_back(o.v);
}

// ---- medium_direct_lighting_inner!, medium-scatter.jl:15-114 -------------------------------------
inline bool medium_direct_lighting(const LightCtx& LC, V3 p, V3 wo, const Wavelengths& lam, const Spec& beta, const Spec& r_u_in, float g,
                                   uint32_t medium, int32_t pixel_index, float time, float light_select, V2 u_light, int32_t num_lights, ShadowWork& out) {
    if (num_lights < 1) return false;
    float pmf;
    int32_t li = bvh_sample_light(*LC.sampler, p, V3(0.0f), light_select, pmf);
    if (li < 1 || li > num_lights || pmf <= 0.0f) return false;
    LightSample ls = sample_light(LC, LC.lights[li - 1], p, lam, u_light);
    if (!(ls.pdf > 0.0f && !is_black(ls.Li))) return false;
    float c = dot(wo, ls.wi);
    float phase = hg_p(g, c);
    if (!(phase > 0.0f)) return false;
    Spec Ld = beta * phase * ls.Li;
    float light_pdf = ls.pdf * pmf;
    float phase_pdf = ls.is_delta ? 0.0f : phase;
    float t_max = ls.is_delta ? norm(ls.p_light - p) - 0.001f : 1.0e6f;
    out.ray = Ray{p, ls.wi, t_max, time}; out.t_max = t_max; out.lambda = lam; out.Ld = Ld;
    out.r_u = r_u_in * phase_pdf; out.r_l = r_u_in * light_pdf; out.pixel_index = pixel_index; out.medium = medium;
    return true;
}
// ---- medium_scatter_inner!, medium-scatter.jl:148-203 ---------------------------------------------
inline bool medium_scatter(V3 p, V3 wo, float time, const Wavelengths& lam, const Spec& beta, const Spec& r_u, float g, uint32_t medium,
                           int32_t depth, int32_t pixel_index, int32_t max_depth, V2 u, RayWork& out) {
    int32_t nd = depth + 1;
    if (nd >= max_depth) return false;
    float pdf;
    V3 wi = sample_hg(g, wo, u, pdf);
    if (!(pdf > 0.0f)) return false;
    out.ray = Ray{p, wi, INF_F, time}; out.depth = nd; out.lambda = lam; out.pixel_index = pixel_index;
    out.beta = beta; out.r_u = r_u; out.r_l = r_u / pdf; out.prev_p = p; out.prev_n = wo; out.eta_scale = 1.0f;
    out.specular_bounce = false; out.any_non_specular = true; out.medium = medium;
    return true;
}

struct Scene {
    // deep copies of everything uploaded
    std::vector<uint32_t> sobol; std::vector<float> cie_x, cie_y, cie_z, d65, rgb_scale, rgb_coeffs;
    Tables T;
    std::vector<float> positions, normals, tangents, uvs; std::vector<uint32_t> indices, tri_meta;
    bool has_normals = false, has_tangents = false, has_uvs = false;
    // bench.py's bounded CPU sample of a full-resolution frame: only image rows y with (y - 1) % row_step == row_offset start a
    // camera ray (same camera, resolution and per-pixel sample streams; the other pixels stay black).  1 / 0 = every row.
    int32_t row_step = 1, row_offset = 0;
    Accel accel;
    InstancedAccel iaccel;       // HkGeometry.instances: per-mesh accelerators + instance transforms (empty for a plain triangle soup)
    std::vector<HkMaterial> materials; std::vector<HkMediumInterface> interfaces;
    std::vector<float> spec_lambdas, spec_values; std::vector<uint32_t> spec_offsets; HkSpectra spectra;
    std::vector<HkLight> lights;
    std::vector<HkEnvMap> envmaps; std::vector<std::vector<float>> env_store;
    std::vector<HkLightBVHNode> lnodes; std::vector<uint32_t> bit_trails; std::vector<int32_t> inf_idx; HkLightSampler sampler;
    std::vector<Medium> media;
    HkCamera camera; HkFilter filter; std::vector<float> f_func, f_mcdf, f_mfunc, f_ccdf;
    HkRenderParams params;
    bool brute_force = false;
    // film state (volpath-state.jl)
    std::vector<float> pixel_L, pixel_rgb, pixel_weight_sum, wavelengths, pdfs, filter_weight;
    struct RaySamples { float direct_uc; V2 direct_u; float indirect_uc; V2 indirect_u; float indirect_rr; };
    std::vector<RaySamples> pixel_samples;
    uint64_t rays_traced = 0;
    std::vector<float> aux_albedo, aux_normal, aux_depth;      // film.albedo / normal / depth, (H, W) column-major (film.jl:410-488)
    std::string err;

    TextureStore textures;
    std::vector<std::vector<float>> tex_alpha;      // per texture: alpha plane (h, w) column-major, empty = opaque (the 4th float of the reference's RGBSpectrum texels)
    // get_surface_alpha (spectral-eval.jl:3882-3888): alpha of the POINT-sampled Kd texel of a MatteMaterial (_sample_texture_data,
    // textures/basic.jl:19-25), 1 for every other material (a MixMaterial included)
    float surface_alpha(uint32_t material, V2 uv) const {
        const HkMaterial& m = materials[material - 1];
        if (m.type != HK_MAT_MATTE || m.tex[0] <= 0 || (size_t)m.tex[0] > tex_alpha.size() || tex_alpha[m.tex[0] - 1].empty()) return 1.0f;
        const int h = textures.h[m.tex[0] - 1], w = textures.w[m.tex[0] - 1];
        int row = (int)(1.0f + (float)(h - 1) * (1.0f - uv.y)), col = (int)(1.0f + (float)(w - 1) * uv.x);      // unsafe_trunc
        row = clampi(row, 1, h); col = clampi(col, 1, w);
        return tex_alpha[m.tex[0] - 1][(size_t)(col - 1) * h + (row - 1)];
    }
    MatCtx matctx() const { MatCtx c; c.T = &T; c.spectra = &spectra; c.textures = &textures; return c; }
    MatCtx matctx_at(V2 uv, uint32_t face_idx, const float* bary) const {                // the TextureFilterContext of one hit
        MatCtx c = matctx(); c.uv = uv; c.face_idx = face_idx; c.bary[0] = bary[0]; c.bary[1] = bary[1]; c.bary[2] = bary[2]; return c;
    }
    LightCtx lightctx() const { return LightCtx{&T, lights.data(), (uint32_t)lights.size(), envmaps.data(), (uint32_t)envmaps.size(), &sampler}; }
    MediaCtx mediactx() const { return MediaCtx{&T, media.data(), (uint32_t)media.size()}; }

    Hit closest_hit(V3 o, V3 d, float t_max) const {
        if (iaccel.enabled()) return iaccel.closest_hit(o, d, t_max, brute_force);
        return brute_force ? accel.closest_hit_brute(o, d, t_max) : accel.closest_hit_bvh(o, d, t_max);
    }
    // A global primitive id resolves to (instance, triangle of the index array) in instanced scenes (HkGeometry.instances): vertices and
    // normals are then taken to world space per hit -- O v and normalize(W^T n), f32, in this operation order (the CUDA path's
    // prim_vertices / prim_normals do the same) -- and TriangleMeta is (instance interface, face + 1, no area light).
    struct PrimRef { const Instance* inst; uint32_t tri; };
    PrimRef resolve(uint32_t prim) const {
        if (!iaccel.enabled()) return PrimRef{nullptr, prim};
        const Instance& I = iaccel.inst[iaccel.instance_of(prim)];
        return PrimRef{&I, I.first_tri + (prim - I.prim_base)};
    }
    struct PrimMeta { uint32_t iface, face, arealight; };
    PrimMeta meta_of(uint32_t prim) const {
        if (!iaccel.enabled()) return PrimMeta{tri_meta[3 * (size_t)prim], tri_meta[3 * (size_t)prim + 1], tri_meta[3 * (size_t)prim + 2]};
        const PrimRef r = resolve(prim);
        return PrimMeta{r.inst->iface, prim - r.inst->prim_base + 1u, 0u};
    }
    V3 vert(uint32_t prim, int k) const {
        const PrimRef r = resolve(prim);
        const float* p = positions.data() + 3 * (size_t)indices[3 * (size_t)r.tri + k];
        return r.inst ? inst_point(r.inst->o2w, V3(p[0], p[1], p[2])) : V3(p[0], p[1], p[2]);
    }
    V3 nrm(uint32_t prim, int k) const {
        if (!has_normals) return V3(NAN, NAN, NAN);
        const PrimRef r = resolve(prim);
        const float* p = normals.data() + 3 * (size_t)indices[3 * (size_t)r.tri + k];
        if (r.inst && !std::isnan(p[0])) return inst_normal(r.inst->w2o, V3(p[0], p[1], p[2]));
        return V3(p[0], p[1], p[2]);
    }
    V3 tan(uint32_t prim, int k) const {
        if (!has_tangents) return V3(NAN, NAN, NAN);
        const PrimRef r = resolve(prim);
        const float* p = tangents.data() + 3 * (size_t)indices[3 * (size_t)r.tri + k];
        if (r.inst && !std::isnan(p[0])) return normalize(inst_vector(r.inst->o2w, V3(p[0], p[1], p[2])));
        return V3(p[0], p[1], p[2]);
    }
    V2 uv(uint32_t prim, int k) const {
        if (!has_uvs) return V2(0, 0);
        const float* p = uvs.data() + 2 * (size_t)indices[3 * (size_t)resolve(prim).tri + k]; return V2(p[0], p[1]);
    }
    // intersection.jl:13-21
    V3 geometric_normal(uint32_t prim) const {
        V3 v0 = vert(prim, 0), v1 = vert(prim, 1), v2 = vert(prim, 2);
        return normalize(cross(v1 - v0, v2 - v0));
    }
    float tri_area(uint32_t prim) const {   // Raycore.area(primitive): half the cross-product norm
        V3 v0 = vert(prim, 0), v1 = vert(prim, 1), v2 = vert(prim, 2);
        return 0.5f * norm(cross(v1 - v0, v2 - v0));
    }
    // intersection.jl:28-37
    V2 uv_bary(uint32_t prim, const float* b) const {
        V2 a = uv(prim, 0), c = uv(prim, 1), d = uv(prim, 2);
        return V2(b[0] * a.x + b[1] * c.x + b[2] * d.x, b[0] * a.y + b[1] * c.y + b[2] * d.y);
    }
    // intersection.jl:158-182 (+ :46-150)
    SurfGeom surface_geometry(uint32_t prim, const float* bary, V3 o, V3 d, float t) const {
        SurfGeom g;
        g.pi = o + d * t;
        V3 n = geometric_normal(prim);
        g.uv = uv_bary(prim, bary);
        // partial derivatives :46-75
        V3 v0 = vert(prim, 0), v1 = vert(prim, 1), v2 = vert(prim, 2);
        V2 uv0 = uv(prim, 0), uv1 = uv(prim, 1), uv2 = uv(prim, 2);
        V2 d10(uv1.x - uv0.x, uv1.y - uv0.y), d20(uv2.x - uv0.x, uv2.y - uv0.y);
        V3 p10 = v1 - v0, p20 = v2 - v0;
        float det = d10.x * d20.y - d10.y * d20.x;
        if (std::fabs(det) < 1.0e-8f) {
            V3 e1 = normalize(p10);
            V3 nn = normalize(cross(p10, p20));
            g.dpdu = e1; g.dpdv = cross(nn, e1);
        } else {
            float inv = 1.0f / det;
            g.dpdu = (d20.y * p10 - d10.y * p20) * inv;
            g.dpdv = (-d20.x * p10 + d10.x * p20) * inv;
        }
        // shading normal :129-150
        V3 n0 = nrm(prim, 0), n1 = nrm(prim, 1), n2 = nrm(prim, 2);
        V3 ns;
        if (std::isnan(n0.x) || std::isnan(n1.x) || std::isnan(n2.x)) ns = n;
        else ns = normalize(V3(bary[0] * n0.x + bary[1] * n1.x + bary[2] * n2.x,
                               bary[0] * n0.y + bary[1] * n1.y + bary[2] * n2.y,
                               bary[0] * n0.z + bary[1] * n1.z + bary[2] * n2.z));
        g.ns = ns;
        g.n = dot(n, ns) < 0.0f ? -n : n;
        // shading tangents :84-122
        V3 t0 = tan(prim, 0), t1 = tan(prim, 1), t2 = tan(prim, 2);
        bool has_t = !std::isnan(t0.x) && !std::isnan(t1.x) && !std::isnan(t2.x);
        V3 dpdus;
        if (has_t) {
            dpdus = normalize(V3(bary[0] * t0.x + bary[1] * t1.x + bary[2] * t2.x,
                                 bary[0] * t0.y + bary[1] * t1.y + bary[2] * t2.y,
                                 bary[0] * t0.z + bary[1] * t1.z + bary[2] * t2.z));
        } else {
            dpdus = g.dpdu - ns * dot(ns, g.dpdu);
            float l2 = dot(dpdus, dpdus);
            if (l2 > 1.0e-10f) dpdus = dpdus / std::sqrt(l2);
            else if (std::fabs(ns.x) > std::fabs(ns.y)) dpdus = V3(-ns.z, 0.0f, ns.x) / std::sqrt(ns.x * ns.x + ns.z * ns.z);
            else dpdus = V3(0.0f, ns.z, -ns.y) / std::sqrt(ns.y * ns.y + ns.z * ns.z);
        }
        g.dpdus = dpdus; g.dpdvs = cross(ns, dpdus);
        return g;
    }

    void alloc_film() {
        size_t n = (size_t)params.width * params.height;
        pixel_L.assign(4 * n, 0.0f); pixel_rgb.assign(3 * n, 0.0f); pixel_weight_sum.assign(n, 0.0f);
        wavelengths.assign(4 * n, 0.0f); pdfs.assign(4 * n, 0.0f); filter_weight.assign(n, 0.0f);
        pixel_samples.assign(n, RaySamples());
        aux_albedo.clear(); aux_normal.clear(); aux_depth.clear();
    }
    // aux_buffer_kernel!, film.jl:433-488: one centre-of-pixel primary ray per pixel, lens sample (0.5, 0.5).  si.core.n / si.core.p are
    // Raycore's (source unavailable): taken as the geometric normal on the shading-normal side and o + d t, as VolPath's own
    // vp_compute_surface_geometry does (intersection.jl:13-21, 158-182).
    void fill_aux_buffers(bool has_infinite_lights) {
        const int W = params.width, H = params.height;
        const size_t n = (size_t)W * H;
        aux_albedo.assign(3 * n, 0.0f); aux_normal.assign(3 * n, 0.0f); aux_depth.assign(n, 0.0f);
        const float miss_depth = has_infinite_lights ? 1.0e30f : INF_F;
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t idx = 0; idx < (int64_t)n; idx++) {
            const int row = (int)(idx % H) + 1, col = (int)(idx / H) + 1;
            Ray ray = camera_generate_ray(camera, V2((float)col + 0.5f, (float)row + 0.5f), V2(0.5f, 0.5f), 0.0f);
            Hit h = closest_hit(ray.o, ray.d, INF_F);
            if (h.hit) {
                const float bary[3] = {1.0f - h.b1 - h.b2, h.b1, h.b2};
                SurfGeom g = surface_geometry(h.prim, bary, ray.o, ray.d, h.t);
                V3 v = g.pi - ray.o;
                aux_normal[3 * idx] = g.n.x; aux_normal[3 * idx + 1] = g.n.y; aux_normal[3 * idx + 2] = g.n.z;
                aux_depth[idx] = std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
                aux_albedo[3 * idx] = aux_albedo[3 * idx + 1] = aux_albedo[3 * idx + 2] = 0.8f;
            } else aux_depth[idx] = miss_depth;
        }
    }
    void accumulate(int32_t pixel_index, const Spec& c) {   // spectral.jl:272-277
        size_t b = (size_t)(pixel_index - 1) * 4;
        for (int i = 0; i < 4; i++) pixel_L[b + i] += c.v[i];
    }

    // ---- detect_camera_medium, intersection.jl:690-747 ----------------------------------------
    uint32_t detect_camera_medium() {
        V3 cam = xform_point(camera.camera_to_world, V3(0.0f));
        V3 d(0.57735027f, 0.57735027f, 0.57735027f);
        V3 o = cam;
        for (int it = 0; it < 16; it++) {
            Hit h = closest_hit(o, d, INF_F);
            if (!h.hit) return 0;
            const HkMediumInterface& mi = interfaces[meta_of(h.prim).iface - 1];
            V3 n = geometric_normal(h.prim);
            if (mi.inside != mi.outside) return dot(-d, n) > 0.0f ? mi.outside : mi.inside;
            V3 pi = o + d * h.t;
            V3 off = dot(d, n) > 0.0f ? n : -n;
            o = pi + off * 1.0e-4f;
        }
        return 0;
    }

    void render_sample(int32_t sample_idx);   // render!, volpath.jl:445-636
    void trace_shadow(const ShadowWork& w);
};

// ---- MixMaterial, src/materials/mix-material.jl ------------------------------------------------------
// mix_hash_float :114-158 (the UInt32 shifts truncate, the SetKey shifts are 64-bit), choose_material :178-196 (`amount` a
// constant or a texture), resolve_mix_material :253-268 (at most 8 levels).
inline float mix_hash_float(V3 p, V3 wo, uint32_t type1, uint32_t vec1, uint32_t type2, uint32_t vec2) {
    auto fb = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    uint64_t h = 0;
    h ^= (uint64_t)fb(p.x);
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)(uint32_t)(fb(p.y) << 4);
    h *= 0x1b873593ull;
    h ^= (uint64_t)(uint32_t)(fb(p.z) << 8);
    h ^= (uint64_t)(uint32_t)(fb(wo.x) << 16);
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)fb(wo.y);
    h *= 0x1b873593ull;
    h ^= (uint64_t)(uint32_t)(fb(wo.z) << 12);
    h ^= (uint64_t)type1 << 24;
    h ^= (uint64_t)vec1;
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)type2 << 28;
    h ^= (uint64_t)vec2 << 4;
    h *= 0x1b873593ull;
    h ^= h >> 31; h *= 0x7fb5d329728ea185ull;
    h ^= h >> 27; h *= 0x81dadef4bc2dd44dull;
    h ^= h >> 33;
    return (float)(uint32_t)(h & 0xFFFFFFFFull) * 2.3283064365386963e-10f;
}
inline uint32_t resolve_mix_material(const std::vector<HkMaterial>& materials, const TextureStore& textures, uint32_t idx, V3 p, V3 wo, V2 uv) {
    for (int it = 0; it < 8; it++) {
        const HkMaterial& m = materials[idx - 1];
        if (m.type != HK_MAT_MIX) return idx;
        float amt = m.f[0];
        if (m.ftex[0] > 0) { float rgb[3]; sample_texture_bilinear(textures, m.ftex[0], uv, rgb); amt = rgb[0]; }      // amt = eval_tex(ctx, mix.amount, uv), :183
        if (amt <= 0.0f) idx = (uint32_t)m.ival[0];
        else if (amt >= 1.0f) idx = (uint32_t)m.ival[1];
        else {
            const float u = mix_hash_float(p, wo, m.flags & 0xFFu, (uint32_t)m.spec[0], (m.flags >> 8) & 0xFFu, (uint32_t)m.spec[1]);
            idx = amt < u ? (uint32_t)m.ival[0] : (uint32_t)m.ival[1];
        }
    }
    return idx;
}

// ---- russian_roulette_spectral, material-dispatch.jl:263-287 ------------------------------------
inline bool russian_roulette(Spec& beta, int32_t depth, float rr) {
    if (depth <= 3) return true;
    float q = std::max(0.05f, 1.0f - max_component(beta));
    if (rr < q) return false;
    beta = beta * (1.0f / (1.0f - q));
    return true;
}

// ---- trace_shadow_transmittance + shadow kernel, intersection.jl:302-406, 565-600 -------------------
inline void Scene::trace_shadow(const ShadowWork& work) {
    Spec T_ray(1.0f), tr_u(1.0f), tr_l(1.0f);
    bool visible = false, done = false;
    uint32_t cur = work.medium;
    V3 o = work.ray.o, dir = work.ray.d;
    float t_rem = work.t_max;
    MediaCtx MC = mediactx();
    for (int it = 0; it < 10 && !done; it++) {
        if (t_rem < 1.0e-6f) break;
        Hit h = closest_hit(o, dir, t_rem);
        #pragma omp atomic
        rays_traced++;
        if (!h.hit) {
            if (cur != 0) {
                Spec sT, su, sl; transmittance_ratio_tracking(MC, cur, o, dir, t_rem, work.lambda, sT, su, sl);
                T_ray = T_ray * sT; tr_u = tr_u * su; tr_l = tr_l * sl;
            }
            visible = true; done = true; break;
        }
        const HkMediumInterface& mi = interfaces[meta_of(h.prim).iface - 1];
        V3 n = geometric_normal(h.prim);
        bool entering = dot(dir, n) < 0.0f;
        bool transmissive = mi.inside != mi.outside;
        if (!transmissive) {
            // stochastic alpha pass-through (:349-372): same hash-seeded test as the trace kernel; the medium is unchanged
            const float bary_a[3] = {1.0f - h.b1 - h.b2, h.b1, h.b2};
            const float alpha = surface_alpha(mi.material, uv_bary(h.prim, bary_a));
            if (alpha < 1.0f) {
                PCG32 arng = pcg32_init(pbrt_hash(o), pbrt_hash(dir));
                if (pcg32_f32(arng) > alpha) {
                    if (cur != 0) {
                        Spec sT, su, sl; transmittance_ratio_tracking(MC, cur, o, dir, h.t, work.lambda, sT, su, sl);
                        T_ray = T_ray * sT; tr_u = tr_u * su; tr_l = tr_l * sl;
                    }
                    o = o + dir * (h.t + 1.0e-4f);
                    t_rem = t_rem - h.t - 1.0e-4f;
                    continue;
                }
            }
            T_ray = Spec(0.0f); tr_u = Spec(1.0f); tr_l = Spec(1.0f); visible = false; done = true; break;
        }
        if (cur != 0) {
            Spec sT, su, sl; transmittance_ratio_tracking(MC, cur, o, dir, h.t, work.lambda, sT, su, sl);
            T_ray = T_ray * sT; tr_u = tr_u * su; tr_l = tr_l * sl;
        }
        if (is_black(T_ray)) { visible = true; done = true; break; }
        cur = entering ? mi.inside : mi.outside;
        o = o + dir * (h.t + 1.0e-4f);
        t_rem = t_rem - h.t - 1.0e-4f;
    }
    if (!done) { T_ray = Spec(0.0f); tr_u = Spec(1.0f); tr_l = Spec(1.0f); visible = false; }
    if (visible && !is_black(T_ray)) {
        Spec mis = work.r_u * tr_u + work.r_l * tr_l;
        float den = average(mis);
        if (den > 1.0e-10f) {
            Spec fin = work.Ld * T_ray / den;
            if (!is_black(fin)) accumulate(work.pixel_index, fin);
        }
    }
}

inline void Scene::render_sample(int32_t sample_idx) {
    const int32_t W = params.width, H = params.height;
    const int64_t n_pixels = (int64_t)W * H;
    SobolRNG rng{T.sobol, params.sobol_log2_spp, params.sobol_n_base4_digits, params.sampler_seed, W};
    MatCtx MC = matctx(); LightCtx LC = lightctx(); MediaCtx MDC = mediactx();
    const int32_t num_lights = (int32_t)lights.size();
    uint32_t initial_medium = detect_camera_medium();
    std::fill(pixel_L.begin(), pixel_L.end(), 0.0f);   // reset_film!

    // ---- vp_generate_camera_rays_kernel!, volpath.jl:125-205 ----------------------------------
    std::vector<Opt<RayWork>> gen((size_t)n_pixels);
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t idx = 1; idx <= n_pixels; idx++) {
        int32_t pixel_idx = (int32_t)(idx - 1);
        int32_t x = pixel_idx % W + 1, y = pixel_idx / W + 1;
        if (row_step > 1 && (y - 1) % row_step != row_offset) {
            filter_weight[idx - 1] = 0.0f;
            for (int i = 0; i < 4; i++) { wavelengths[(size_t)pixel_idx * 4 + i] = 538.0f; pdfs[(size_t)pixel_idx * 4 + i] = 1.0f; }
            continue;
        }
        float wavelength_u = zsobol_1d(rng, x, y, sample_idx, 1);
        V2 jitter = zsobol_2d(rng, x, y, sample_idx, 3);
        float time_u = zsobol_1d(rng, x, y, sample_idx, 4);
        V2 lens = zsobol_2d(rng, x, y, sample_idx, 6);
        FilterSample fs = filter_sample(filter, jitter);
        filter_weight[idx - 1] = fs.weight;
        Wavelengths lam = sample_wavelengths_visible(wavelength_u);
        for (int i = 0; i < 4; i++) { wavelengths[(size_t)pixel_idx * 4 + i] = lam.lambda[i]; pdfs[(size_t)pixel_idx * 4 + i] = lam.pdf[i]; }
        V2 p_film((float)x + 0.5f + fs.p.x, (float)H - (float)y + 1.0f + 0.5f + fs.p.y);
        Ray ray = camera_generate_ray(camera, p_film, lens, time_u);
        RayWork w;
        w.ray = ray; w.ray.time = 0.0f;   // Raycore.Ray(o, d, t_max) drops the camera time (volpath.jl:184)
        w.depth = 0; w.lambda = lam; w.pixel_index = (int32_t)idx;
        w.beta = Spec(1.0f); w.r_u = Spec(1.0f); w.r_l = Spec(1.0f);
        w.prev_p = V3(0.0f); w.prev_n = V3(0, 0, 1); w.eta_scale = 1.0f;
        w.specular_bounce = false; w.any_non_specular = false; w.medium = initial_medium;
        gen[idx - 1].valid = true; gen[idx - 1].v = w;
    }
    std::vector<RayWork> ray_queue, next_queue;
    compact(gen, ray_queue);
    gen.clear(); gen.shrink_to_fit();

    for (int32_t depth = 0; depth < params.max_depth; depth++) {
        const int64_t n_rays = (int64_t)ray_queue.size();
        if (n_rays == 0) break;
        // ---- vp_generate_ray_samples_kernel!, volpath.jl:222-271 -------------------------------
        #pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < n_rays; i++) {
            int32_t pi = ray_queue[i].pixel_index;
            int32_t p0 = pi - 1;
            int32_t px = p0 % W + 1, py = p0 / W + 1;
            int32_t base = 6 + 7 * depth;
            RaySamples s;
            s.direct_uc = zsobol_1d(rng, px, py, sample_idx, base + 1);
            s.direct_u = zsobol_2d(rng, px, py, sample_idx, base + 3);
            s.indirect_uc = zsobol_1d(rng, px, py, sample_idx, base + 4);
            s.indirect_u = zsobol_2d(rng, px, py, sample_idx, base + 6);
            s.indirect_rr = zsobol_1d(rng, px, py, sample_idx, base + 7);
            pixel_samples[pi - 1] = s;
        }
        // ---- vp_trace_rays_kernel!, intersection.jl:188-269 ------------------------------------
        std::vector<Opt<MediumSampleWork>> o_med((size_t)n_rays);
        std::vector<Opt<EscapedWork>> o_esc((size_t)n_rays);
        std::vector<Opt<HitSurfaceWork>> o_hit((size_t)n_rays);
        #pragma omp parallel for schedule(dynamic, 64)
        for (int64_t i = 0; i < n_rays; i++) {
            const RayWork& w = ray_queue[i];
            Hit h = closest_hit(w.ray.o, w.ray.d, w.ray.t_max);
            #pragma omp atomic
            rays_traced++;
            if (w.medium != 0) {
                MediumSampleWork m; m.w = w;
                if (h.hit) {
                    const PrimMeta pm = meta_of(h.prim); const uint32_t meta[3] = {pm.iface, pm.face, pm.arealight};
                    const HkMediumInterface& mi = interfaces[meta[0] - 1];
                    float bary[3] = {1.0f - h.b1 - h.b2, h.b1, h.b2};
                    m.t_max = h.t; m.has_surface_hit = true;
                    m.g = surface_geometry(h.prim, bary, w.ray.o, w.ray.d, h.t);
                    m.material = mi.material; m.iface = mi; m.face_idx = meta[1];
                    m.bary[0] = bary[0]; m.bary[1] = bary[1]; m.bary[2] = bary[2];
                    m.arealight_flat_idx = meta[2]; m.triangle_area = tri_area(h.prim);
                } else {
                    m.t_max = INF_F; m.has_surface_hit = false; m.g = SurfGeom(); m.g.n = V3(0, 0, 1); m.g.ns = V3(0, 0, 1);
                    m.material = 0; m.iface = HkMediumInterface{0, 0, 0}; m.face_idx = 0; m.bary[0] = m.bary[1] = m.bary[2] = 0.0f;
                    m.arealight_flat_idx = 0; m.triangle_area = 0.0f;
                }
                o_med[i].valid = true; o_med[i].v = m;
                continue;
            }
            // vacuum: the alpha loop (:221-266).  Alpha-killed surfaces are skipped without consuming depth: the ray restarts 1e-4 behind
            // the surface, at most 16 times; a ray that is still being skipped after that is absorbed.
            Ray ray = w.ray;
            bool absorbed = false;
            for (int pass = 0; h.hit; pass++) {
                const HkMediumInterface& mi_a = interfaces[meta_of(h.prim).iface - 1];
                const float bary_a[3] = {1.0f - h.b1 - h.b2, h.b1, h.b2};
                const float alpha = surface_alpha(mi_a.material, uv_bary(h.prim, bary_a));
                if (!(alpha < 1.0f)) break;
                PCG32 arng = pcg32_init(pbrt_hash(ray.o), pbrt_hash(ray.d));
                if (!(pcg32_f32(arng) > alpha)) break;                           // kept: a regular surface hit
                if (pass == 15) { absorbed = true; break; }                      // the 16th iteration also asked to continue
                V3 pi_a = ray.o + ray.d * h.t;
                V3 n_a = geometric_normal(h.prim);
                V3 off = dot(ray.d, n_a) > 0.0f ? n_a : -n_a;
                ray = Ray{pi_a + off * 1.0e-4f, ray.d, INF_F, 0.0f};
                h = closest_hit(ray.o, ray.d, ray.t_max);
                #pragma omp atomic
                rays_traced++;
            }
            if (absorbed) continue;
            if (!h.hit) {
                EscapedWork e{w.ray.d, w.lambda, w.pixel_index, w.beta, w.r_u, w.r_l, w.depth, w.specular_bounce, w.prev_p, w.prev_n};
                o_esc[i].valid = true; o_esc[i].v = e;
                continue;
            }
            const PrimMeta pm = meta_of(h.prim); const uint32_t meta[3] = {pm.iface, pm.face, pm.arealight};
            const HkMediumInterface& mi = interfaces[meta[0] - 1];
            float bary[3] = {1.0f - h.b1 - h.b2, h.b1, h.b2};
            HitSurfaceWork hs;
            hs.ray = w.ray; hs.g = surface_geometry(h.prim, bary, ray.o, ray.d, h.t);
            hs.material = mi.material; hs.iface = mi; hs.face_idx = meta[1];
            hs.bary[0] = bary[0]; hs.bary[1] = bary[1]; hs.bary[2] = bary[2];
            hs.arealight_flat_idx = meta[2]; hs.triangle_area = tri_area(h.prim);
            hs.lambda = w.lambda; hs.pixel_index = w.pixel_index; hs.beta = w.beta; hs.r_u = w.r_u; hs.r_l = w.r_l;
            hs.depth = w.depth; hs.eta_scale = w.eta_scale; hs.specular_bounce = w.specular_bounce; hs.any_non_specular = w.any_non_specular;
            hs.prev_p = w.prev_p; hs.prev_n = w.prev_n; hs.current_medium = w.medium; hs.t_hit = h.t;
            o_hit[i].valid = true; o_hit[i].v = hs;
        }
        std::vector<MediumSampleWork> medium_sample_queue; compact(o_med, medium_sample_queue);
        std::vector<EscapedWork> escaped_queue; compact(o_esc, escaped_queue);
        std::vector<HitSurfaceWork> hit_surface_queue; compact(o_hit, hit_surface_queue);
        o_med.clear(); o_esc.clear(); o_hit.clear();
        std::vector<MediumScatterWork> medium_scatter_queue;
        std::vector<ShadowWork> shadow_queue;
        next_queue.clear();

        // ---- media: delta tracking + medium NEE + phase sampling (volpath.jl:549-564) ------------
        if (!media.empty() && !medium_sample_queue.empty()) {
            const int64_t nm = (int64_t)medium_sample_queue.size();
            std::vector<Opt<MediumScatterWork>> o_sc((size_t)nm);
            std::vector<Opt<HitSurfaceWork>> o_h2((size_t)nm);
            std::vector<Opt<EscapedWork>> o_e2((size_t)nm);
            #pragma omp parallel for schedule(dynamic, 16)
            for (int64_t i = 0; i < nm; i++) {
                const MediumSampleWork& m = medium_sample_queue[i];
                Spec Ladd(0.0f);
                DeltaResult r = sample_medium_interaction(MDC, m.w.medium, m.w.ray.o, m.w.ray.d, m.t_max, m.w.lambda, m.w.beta, m.w.r_u, m.w.r_l, m.w.depth, params.max_depth, &Ladd);
                if (!is_black(Ladd)) accumulate(m.w.pixel_index, Ladd);
                if (r.event == DeltaResult::SCATTER) {
                    MediumScatterWork s{r.p, -m.w.ray.d, m.w.ray.time, m.w.lambda, m.w.pixel_index, r.beta, r.r_u, m.w.depth, m.w.medium, r.g};
                    o_sc[i].valid = true; o_sc[i].v = s;
                } else if (r.event == DeltaResult::SURVIVED) {
                    if (m.has_surface_hit) {
                        HitSurfaceWork hs;
                        hs.ray = m.w.ray; hs.g = m.g; hs.material = m.material; hs.iface = m.iface; hs.face_idx = m.face_idx;
                        hs.bary[0] = m.bary[0]; hs.bary[1] = m.bary[1]; hs.bary[2] = m.bary[2];
                        hs.arealight_flat_idx = m.arealight_flat_idx; hs.triangle_area = m.triangle_area;
                        hs.lambda = m.w.lambda; hs.pixel_index = m.w.pixel_index; hs.beta = r.beta; hs.r_u = r.r_u; hs.r_l = r.r_l;
                        hs.depth = m.w.depth; hs.eta_scale = m.w.eta_scale; hs.specular_bounce = m.w.specular_bounce; hs.any_non_specular = m.w.any_non_specular;
                        hs.prev_p = m.w.prev_p; hs.prev_n = m.w.prev_n; hs.current_medium = m.w.medium; hs.t_hit = m.t_max;
                        o_h2[i].valid = true; o_h2[i].v = hs;
                    } else {
                        EscapedWork e{m.w.ray.d, m.w.lambda, m.w.pixel_index, r.beta, r.r_u, r.r_l, m.w.depth, m.w.specular_bounce, m.w.prev_p, m.w.prev_n};
                        o_e2[i].valid = true; o_e2[i].v = e;
                    }
                }   // ABSORBED: path ends
            }
            compact(o_sc, medium_scatter_queue); compact(o_h2, hit_surface_queue); compact(o_e2, escaped_queue);
        }
        if (!media.empty() && !medium_scatter_queue.empty()) {
            const int64_t ns = (int64_t)medium_scatter_queue.size();
            std::vector<Opt<ShadowWork>> o_sh((size_t)ns);
            std::vector<Opt<RayWork>> o_r((size_t)ns);
            #pragma omp parallel for schedule(dynamic, 64)
            for (int64_t i = 0; i < ns; i++) {
                const MediumScatterWork& s = medium_scatter_queue[i];
                const RaySamples& smp = pixel_samples[s.pixel_index - 1];
                if (num_lights > 0) {
                    ShadowWork sh;
                    if (medium_direct_lighting(LC, s.p, s.wo, s.lambda, s.beta, s.r_u, s.g, s.medium, s.pixel_index, s.time, smp.direct_uc, smp.direct_u, num_lights, sh)) {
                        o_sh[i].valid = true; o_sh[i].v = sh;
                    }
                }
                RayWork nr;
                if (medium_scatter(s.p, s.wo, s.time, s.lambda, s.beta, s.r_u, s.g, s.medium, s.depth, s.pixel_index, params.max_depth, smp.indirect_u, nr)) {
                    o_r[i].valid = true; o_r[i].v = nr;
                }
            }
            compact(o_sh, shadow_queue); compact(o_r, next_queue);
        }

        // ---- vp_handle_escaped_rays_kernel!, intersection.jl:622-668 -----------------------------
        if (!escaped_queue.empty() && num_lights > 0) {
            const int64_t ne = (int64_t)escaped_queue.size();
            #pragma omp parallel for schedule(dynamic, 256)
            for (int64_t i = 0; i < ne; i++) {
                const EscapedWork& w = escaped_queue[i];
                Spec Le = evaluate_escaped_ray(LC, w.ray_d, w.lambda);
                Spec contrib = w.beta * Le;
                if (is_black(contrib)) continue;
                Spec fin;
                if (w.depth == 0 || w.specular_bounce) fin = contrib / average(w.r_u);
                else {
                    float lcp = num_lights > 0 ? 1.0f / (float)num_lights : 0.0f;
                    float lpdf = compute_env_light_pdf(LC, w.ray_d);
                    Spec r_l = w.r_l * lcp * lpdf;
                    float den = average(w.r_u + r_l);
                    fin = den > 1.0e-10f ? contrib / den : contrib / average(w.r_u);
                }
                accumulate(w.pixel_index, fin);
            }
        }

        const int64_t n_hits = (int64_t)hit_surface_queue.size();
        if (n_hits > 0) {
            // ---- vp_process_surface_hits_kernel!, surface-eval.jl:147-220 -----------------------
            std::vector<MaterialEvalWork> material_queue((size_t)n_hits);
            #pragma omp parallel for schedule(dynamic, 256)
            for (int64_t i = 0; i < n_hits; i++) {
                const HitSurfaceWork& w = hit_surface_queue[i];
                V3 wo = -w.ray.d;
                uint32_t material_idx = resolve_mix_material(materials, textures, w.material, w.g.pi, wo, w.g.uv);   // mix-material.jl:253-268
                if (w.arealight_flat_idx > 0) {
                    const HkLight& L = lights[w.arealight_flat_idx - 1];
                    Spec Le = arealight_Le(LC, L, wo, w.g.n, w.lambda);
                    if (!is_black(Le)) {
                        Spec contrib = w.beta * Le;
                        Spec fin;
                        if (w.depth == 0 || w.specular_bounce) fin = contrib / average(w.r_u);
                        else {
                            float lcp = bvh_pmf(sampler, w.g.pi, w.g.n, (int32_t)w.arealight_flat_idx);
                            float ct = std::fabs(dot(w.g.n, normalize(w.ray.d)));
                            float lightPDF = (ct > 0.0f && w.triangle_area > 0.0f) ? lcp * ((w.t_hit * w.t_hit) / (ct * w.triangle_area)) : 0.0f;
                            Spec r_l = w.r_l * lightPDF;
                            float den = average(w.r_u + r_l);
                            fin = den > 1.0e-10f ? contrib / den : contrib / average(w.r_u);
                        }
                        accumulate(w.pixel_index, fin);
                    }
                }
                MaterialEvalWork m;
                m.g = w.g; m.time = w.ray.time; m.face_idx = w.face_idx; m.bary[0] = w.bary[0]; m.bary[1] = w.bary[1]; m.bary[2] = w.bary[2];
                m.wo = wo; m.material = material_idx; m.iface = w.iface; m.lambda = w.lambda; m.pixel_index = w.pixel_index;
                m.beta = w.beta; m.r_u = w.r_u; m.r_l = w.r_l; m.depth = w.depth; m.eta_scale = w.eta_scale;
                m.specular_bounce = w.specular_bounce; m.any_non_specular = w.any_non_specular; m.prev_p = w.prev_p; m.prev_n = w.prev_n;
                m.current_medium = w.current_medium;
                material_queue[i] = m;
            }
            // ---- surface_direct_lighting_inner!, surface-eval.jl:250-342 -------------------------
            if (num_lights > 0) {
                std::vector<Opt<ShadowWork>> o_sh((size_t)n_hits);
                #pragma omp parallel for schedule(dynamic, 64)
                for (int64_t i = 0; i < n_hits; i++) {
                    const MaterialEvalWork& w = material_queue[i];
                    const RaySamples& smp = pixel_samples[w.pixel_index - 1];
                    float pmf;
                    int32_t li = bvh_sample_light(sampler, w.g.pi, w.g.ns, smp.direct_uc, pmf);
                    if (li < 1 || li > num_lights || pmf <= 0.0f) continue;
                    LightSample ls = sample_light(LC, lights[li - 1], w.g.pi, w.lambda, smp.direct_u);
                    if (!(ls.pdf > 0.0f && !is_black(ls.Li))) continue;
                    BSDFEval be = eval_material(matctx_at(w.g.uv, w.face_idx, w.bary), materials[w.material - 1], w.wo, ls.wi, w.g.ns, w.lambda);
                    if (is_black(be.f)) continue;
                    // compute_direct_lighting_spectral, lights.jl:535-600
                    float ct = std::fabs(dot(ls.wi, w.g.ns));
                    Spec Ld = w.beta * be.f * ls.Li * ct;
                    if (is_black(Ld)) continue;
                    V3 off = 1.0e-4f * w.g.ns;
                    V3 ro = dot(ls.wi, w.g.ns) > 0.0f ? (w.g.pi + off) : (w.g.pi - off);
                    V3 tl = ls.p_light - ro;
                    float t_max = std::sqrt(dot(tl, tl)) - 1.0e-3f;
                    float bp = ls.is_delta ? 0.0f : be.pdf;
                    Spec nr_u = w.r_u * bp;
                    Spec nr_l = w.r_u * ls.pdf;
                    ShadowWork sh;
                    sh.ray = Ray{ro, ls.wi, t_max, 0.0f}; sh.t_max = t_max; sh.lambda = w.lambda; sh.Ld = Ld;
                    sh.r_u = nr_u; sh.r_l = nr_l * pmf; sh.pixel_index = w.pixel_index; sh.medium = w.current_medium;
                    o_sh[i].valid = true; o_sh[i].v = sh;
                }
                compact(o_sh, shadow_queue);
            }
            // ---- vp_trace_shadow_rays!, intersection.jl:565-616 (both surface + medium NEE rays) ----
            {
                const int64_t nsq = (int64_t)shadow_queue.size();
                #pragma omp parallel for schedule(dynamic, 64)
                for (int64_t i = 0; i < nsq; i++) trace_shadow(shadow_queue[i]);
                shadow_queue.clear();
            }
            // ---- evaluate_material_inner!, surface-eval.jl:396-512 ------------------------------
            std::vector<Opt<RayWork>> o_r((size_t)n_hits);
            #pragma omp parallel for schedule(dynamic, 64)
            for (int64_t i = 0; i < n_hits; i++) {
                const MaterialEvalWork& w = material_queue[i];
                int32_t new_depth = w.depth + 1;
                if (new_depth >= params.max_depth) continue;
                const RaySamples& smp = pixel_samples[w.pixel_index - 1];
                bool regularize = params.regularize && w.any_non_specular;
                BSDFSample s = sample_material(matctx_at(w.g.uv, w.face_idx, w.bary), materials[w.material - 1], w.wo, w.g.ns, w.lambda, smp.indirect_u, smp.indirect_uc, regularize);
                if (!(s.pdf > 0.0f && !is_black(s.f))) continue;
                float ct = std::fabs(dot(s.wi, w.g.ns));
                Spec nb = s.is_specular ? w.beta * s.f : w.beta * s.f * ct / s.pdf;
                float nes = w.eta_scale * s.eta_scale;
                Spec nrl = s.is_specular ? w.r_u : w.r_u / s.pdf;
                if (!russian_roulette(nb, new_depth, smp.indirect_rr)) continue;
                uint32_t new_medium = (w.iface.inside != w.iface.outside) ? (dot(s.wi, w.g.n) > 0.0f ? w.iface.outside : w.iface.inside) : w.current_medium;
                V3 od = dot(s.wi, w.g.n) > 0.0f ? w.g.n : -w.g.n;
                RayWork r;
                r.ray = Ray{w.g.pi + od * 0.0001f, s.wi, INF_F, 0.0f};
                r.depth = new_depth; r.lambda = w.lambda; r.pixel_index = w.pixel_index; r.beta = nb; r.r_u = w.r_u; r.r_l = nrl;
                r.prev_p = w.g.pi; r.prev_n = w.g.ns; r.eta_scale = nes; r.specular_bounce = s.is_specular;
                r.any_non_specular = w.any_non_specular || !s.is_specular; r.medium = new_medium;
                o_r[i].valid = true; o_r[i].v = r;
            }
            compact(o_r, next_queue);
        } else if (!shadow_queue.empty()) {
            // NOTE (reference quirk, volpath.jl:571-609): shadow rays are only traced inside the `n_hits > 0`
            // branch; medium NEE rays queued in a bounce with zero surface hits are dropped by the next
            // reset_iteration_queues!.  Restated literally.
            shadow_queue.clear();
        }
        ray_queue.swap(next_queue);
    }

    // ---- vp_accumulate_to_rgb_kernel!, volpath.jl:326-375 ------------------------------------
    #pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_pixels; p++) {
        Spec L(pixel_L[4 * p], pixel_L[4 * p + 1], pixel_L[4 * p + 2], pixel_L[4 * p + 3]);
        Wavelengths lam;
        for (int i = 0; i < 4; i++) { lam.lambda[i] = wavelengths[4 * p + i]; lam.pdf[i] = pdfs[4 * p + i]; }
        V3 xyz = spectral_to_xyz(T, L, lam);
        V3 rgb = xyz_to_linear_srgb(xyz);
        rgb = V3(std::max(0.0f, rgb.x), std::max(0.0f, rgb.y), std::max(0.0f, rgb.z));
        float m = std::max(std::max(rgb.x, rgb.y), rgb.z);
        if (m > params.max_component_value) rgb = rgb * (params.max_component_value / m);
        float w = filter_weight[p];
        pixel_rgb[3 * p] += w * rgb.x; pixel_rgb[3 * p + 1] += w * rgb.y; pixel_rgb[3 * p + 2] += w * rgb.z;
        pixel_weight_sum[p] += w;
    }
}

}  // namespace ok
