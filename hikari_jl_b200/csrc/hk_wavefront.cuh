// hk_wavefront.cuh — device-resident scene + SoA path state + queues + the wavefront stage kernels.
//
// What the reference does per bounce with ~12 KernelAbstractions launches over AOS records of 112-340 bytes and a
// device->host `length(queue)` copy per stage (src/integrators/volpath/volpath.jl:538-612, workqueue.jl:108-121) is
// done here with: one SoA path-state pool indexed by slot (= sample-in-batch * n_pixels + pixel), queues of 4-byte
// slot ids, warp- / block-aggregated appends (__ballot_sync / __match_any_sync / __popc, shared-memory counters in the
// routing kernel), per-material queues and kernels, device-side counts (no host round trips inside a sample pass), and
// persistent kernels (traversal, delta / ratio tracking) whose lanes pull new work from a global cursor as they finish.
// Kernels of one bounce:  k_reset_bounce -> k_trace -> k_route -> [k_medium_track -> k_medium_finish]
//                         -> { k_escaped | k_shade<TYPE> ... }  (disjoint slots, forked over side streams by hk_api.cu)
//                         -> k_shadow_opaque  or  { k_shadow_seg_trace -> k_shadow_seg_ratio } x rounds
// once per pass: k_camera before, k_film_accumulate after; once per upload: k_sobol_prefix, k_precompute_uplifts,
// k_patch_tri_types.
#pragma once
#include "hk_math.cuh"
#include "hk_spectral.cuh"
#include "hk_bsdf.cuh"
#include "hk_lights.cuh"
#include "hk_media.cuh"
#include "hk_traverse.cuh"

#define HK_MAX_MAT_TYPES 8
// counter slots
#define HK_C_RAY0 0
#define HK_C_RAY1 1
#define HK_C_ESCAPED 2
#define HK_C_MEDIUM 3
#define HK_C_SHADOW 4
#define HK_C_TOTAL_HITS 5
#define HK_C_CURSOR_TRACE 6
#define HK_C_CURSOR_SHADOW 7
#define HK_C_HIT0 8            // + hit queue 0..7 of the material type: types 1..7 use queue = type, CoatedConductor (9) the spare queue 0
#define HK_C_HIT1 56           // + (hit queue - 8): second bank, CoatedDiffuseTransmission (10) = queue 8
#define HK_N_HIT_QUEUES 10
// shading class = the material type, except a MatteMaterial with a textured Kd, which gets its own class, queue and k_shade
// instantiation (the constant-parameter kernels stay free of texture code)
#define HK_SHADE_MATTE_TEX 11
#define HK_TYPE_QUEUE(t) ((t) == HK_MAT_COATED_CONDUCTOR ? 0 : (t) == HK_MAT_COATED_DIFFUSE_TRANSMISSION ? 8 : (t) == HK_SHADE_MATTE_TEX ? 9 : (t))
#define HK_HIT_COUNTER(q) ((q) < 8 ? HK_C_HIT0 + (q) : HK_C_HIT1 + (q) - 8)
#define HK_N_QUEUE_COUNTERS 16 // the counters above (what hk_bounce_profile reports)
#define HK_C_CURSOR_MEDIUM 16  // k_medium_track work cursor
// second copies of the shadow-pass counters: in opaque-only scenes the shadow pass of bounce b runs on its own stream while
// bounce b+1 is already being traced and routed, so bounces alternate between the two sets (parity = bounce & 1)
#define HK_C_SHADOW_B 17
#define HK_C_CURSOR_SHADOW_B 18
#define HK_C_TOTAL_HITS_B 19
#define HK_CI_SHADOW(par) ((par) ? HK_C_SHADOW_B : HK_C_SHADOW)
#define HK_CI_CURSOR_SHADOW(par) ((par) ? HK_C_CURSOR_SHADOW_B : HK_C_CURSOR_SHADOW)
#define HK_CI_TOTAL_HITS(par) ((par) ? HK_C_TOTAL_HITS_B : HK_C_TOTAL_HITS)
#define HK_C_SHROUND0 20       // + r (1..10): shadow rays still unresolved after r medium-boundary crossings (round 0 = HK_C_SHADOW)
#define HK_C_SHCUR_TRACE 32    // + r: work cursor of the closest-hit pass of shadow round r
#define HK_C_SHCUR_RATIO 44    // + r: work cursor of the ratio-tracking pass of shadow round r
#define HK_SHADOW_ROUNDS 10    // trace_shadow_transmittance: at most 10 segments (intersection.jl:302-406)
#define HK_N_COUNTERS 128
// alpha-tested surfaces (intersection.jl:221-266): round r = 1..15 of the trace stage re-traces the rays whose hit was skipped in
// round r-1 (round 0 = the bounce's own trace); the two retrace queues ping-pong by round parity
#define HK_ALPHA_ROUNDS 16
#define HK_C_ALPHA_N0 64       // + r: rays queued for retrace round r
#define HK_C_ALPHA_CUR0 80     // + r: work cursor of retrace round r
#define HK_C_LBVH 96           // hits of the bounce whose light sample falls on the light BVH (k_hit_lights -> k_hit_lights_bvh)
#ifndef HK_LIGHTS_COMPACT
#define HK_LIGHTS_COMPACT 1
#endif
static_assert(HK_C_SHROUND0 + HK_SHADOW_ROUNDS < HK_C_SHCUR_TRACE && HK_C_SHCUR_TRACE + HK_SHADOW_ROUNDS < HK_C_SHCUR_RATIO &&
              HK_C_SHCUR_RATIO + HK_SHADOW_ROUNDS < HK_C_HIT1, "shadow-round counters overlap the second bank of hit-queue counters");
static_assert(HK_C_ALPHA_N0 + HK_ALPHA_ROUNDS <= HK_C_ALPHA_CUR0 && HK_C_ALPHA_CUR0 + HK_ALPHA_ROUNDS <= HK_N_COUNTERS && HK_C_ALPHA_N0 > HK_C_HIT1 + 1, "alpha-round counters");
static_assert(HK_HIT_COUNTER(7) == 15 && HK_HIT_COUNTER(8) == HK_C_HIT1 && HK_HIT_COUNTER(HK_N_HIT_QUEUES - 1) < HK_N_COUNTERS,
              "hit-queue counters must stay inside the counter block");
static_assert(HK_TYPE_QUEUE(HK_MAT_COATED_CONDUCTOR) == 0 && HK_TYPE_QUEUE(HK_MAT_COATED_DIFFUSE_TRANSMISSION) == 8 &&
              HK_TYPE_QUEUE(HK_SHADE_MATTE_TEX) == 9 && HK_TYPE_QUEUE(HK_MAT_DIFFUSE_TRANSMISSION) == 7 && HK_SHADE_MATTE_TEX < 16,
              "every shading class needs its own hit queue and must fit the 4-bit type field of the hit record");

struct DevScene {
    DevTables T;
    DevBvh bvh;
    // shading geometry (world space)
    const float* __restrict__ positions; const float* __restrict__ normals; const uint32_t* __restrict__ indices; const uint32_t* __restrict__ tri_meta;
    const HkMaterial* __restrict__ materials; const HkMediumInterface* __restrict__ interfaces;
    const float* __restrict__ spec_lambdas; const float* __restrict__ spec_values; const uint32_t* __restrict__ spec_offsets;
    const HkLight* __restrict__ lights; int32_t n_lights;
    const DevEnvMap* __restrict__ envmaps;
    const DevLNode* __restrict__ lnodes; const uint32_t* __restrict__ bit_trails; const int32_t* __restrict__ inf_idx; int32_t n_infinite, n_bvh;
    const int32_t* __restrict__ esc_idx; int32_t n_esc;
    const DevMedium* __restrict__ media; int32_t n_media;
    int32_t any_medium_transition;     // some interface has inside != outside
    int32_t split_lights;              // 1: k_hit_lights does emissive-hit MIS + the NEE light sample ahead of k_shade<TYPE, true> (large light sets)
    HkCamera camera;
    DevFilter filter;
    int32_t width, height, max_depth, regularize;
    float max_component_value;
    SobolParams sobol;
    // textured parameters (appended last: the members above keep their offsets in the kernel parameter block)
    const float* __restrict__ uvs; const HkTexture* __restrict__ textures; int32_t n_textures;
    // instancing (HkGeometry.instances): instances in upload order + their first global primitive ids (ascending); n_inst = 0: plain soup
    const struct DevInstance* __restrict__ instances; const uint32_t* __restrict__ inst_prim_base; int32_t n_inst;
    int32_t has_alpha;                 // some uploaded texture carries an alpha plane: the trace / shadow stages run their pass-through rounds
    int32_t smem_mask_medium; uint32_t smem_mask_words;      // medium (1-based, 0 = none) whose empty-cell mask the tracking kernels stage in shared memory
    uint32_t tex_classes;              // bit c: some material of shading class c has textured parameters (its k_shade runs the TEX instantiation)
};
struct DevInstance { float o2w[12]; float w2o[12]; uint32_t first_tri, prim_base, iface, n_tris; };
// A global primitive id resolves to (instance, triangle of the index array); vertices / normals of an instanced triangle are taken
// to world space per hit: O v and normalize(W^T n) in f32, in the operation order of the oracle's Scene::vert / nrm.
struct PrimRef { uint32_t tri; const DevInstance* inst; };
HK_DEV PrimRef resolve_prim(const DevScene& D, uint32_t prim0) {
    PrimRef r; r.tri = prim0; r.inst = nullptr;
    if (D.n_inst <= 0) return r;
    int lo = 0, hi = D.n_inst - 1;                     // last instance whose first primitive id is <= prim0
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(D.inst_prim_base + mid) <= prim0) lo = mid; else hi = mid - 1; }
    r.inst = D.instances + lo;
    r.tri = r.inst->first_tri + (prim0 - r.inst->prim_base);
    return r;
}
HK_DEV uint32_t prim_iface(const DevScene& D, const PrimRef& r, uint32_t prim0) { return r.inst ? r.inst->iface : __ldg(D.tri_meta + 3 * (size_t)prim0); }
HK_DEV uint32_t prim_iface(const DevScene& D, uint32_t prim0) { return prim_iface(D, resolve_prim(D, prim0), prim0); }
HK_DEV uint32_t prim_arealight(const DevScene& D, const PrimRef& r, uint32_t prim0) { return r.inst ? 0u : __ldg(D.tri_meta + 3 * (size_t)prim0 + 2); }
HK_DEV float3 inst_point(const float* m, float3 p) { return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]); }
HK_DEV float3 inst_normal(const float* w, float3 n) { return norm3(f3(w[0] * n.x + w[4] * n.y + w[8] * n.z, w[1] * n.x + w[5] * n.y + w[9] * n.z, w[2] * n.x + w[6] * n.y + w[10] * n.z)); }
HK_DEV void prim_vertices(const DevScene& D, const PrimRef& r, float3& v0, float3& v1, float3& v2) {
    const uint32_t i0 = __ldg(D.indices + 3 * (size_t)r.tri), i1 = __ldg(D.indices + 3 * (size_t)r.tri + 1), i2 = __ldg(D.indices + 3 * (size_t)r.tri + 2);
    const float* P = D.positions;
    v0 = f3(__ldg(P + 3 * (size_t)i0), __ldg(P + 3 * (size_t)i0 + 1), __ldg(P + 3 * (size_t)i0 + 2));
    v1 = f3(__ldg(P + 3 * (size_t)i1), __ldg(P + 3 * (size_t)i1 + 1), __ldg(P + 3 * (size_t)i1 + 2));
    v2 = f3(__ldg(P + 3 * (size_t)i2), __ldg(P + 3 * (size_t)i2 + 1), __ldg(P + 3 * (size_t)i2 + 2));
    if (r.inst) { v0 = inst_point(r.inst->o2w, v0); v1 = inst_point(r.inst->o2w, v1); v2 = inst_point(r.inst->o2w, v2); }
}
HK_DEV uint32_t shade_class(const HkMaterial& m) { return (m.type == HK_MAT_MATTE && m.tex[0] > 0) ? (uint32_t)HK_SHADE_MATTE_TEX : (uint32_t)m.type; }
struct PathState {
    float4 *ray_a, *ray_b, *hit, *lambda, *lpdf, *beta, *r_u, *r_l, *L;
    uint32_t* flags; float* fweight;
    float4 *sh_a, *sh_b, *sh_Ld, *sh_ru, *sh_rl; uint32_t* sh_medium;
    float4 *med; uint32_t* med_ev;                 // delta-tracking result per slot: (scatter point, g), event
    uint32_t* res_mat;                             // material a MixMaterial hit resolved to (written by the routing, read by k_shade)
    float4 *sh_hit, *sh_T, *sh_tu, *sh_tl;         // shadow rays through media: segment hit, running transmittance / MIS ratios
    float4 *nee_a, *nee_b, *nee_c;                 // light sample of a surface hit, written by k_hit_lights for k_shade: Li | wi, pdf | p_light, pmf (sign bit = delta light)
    uint32_t *q_ray[2], *q_escaped, *q_medium, *q_shadow, *q_shadow2, *q_hit[HK_N_HIT_QUEUES], *q_alpha[2], *q_lbvh;
    uint32_t* counts;                  // [HK_N_COUNTERS]
    unsigned long long* rays_traced;
    unsigned long long* path_vertices;  // surface hits routed + medium scatter events (HkStats::path_vertices)
    float *pixel_rgb, *pixel_weight;   // film accumulators
};
struct PassArgs { int32_t first_sample, stride, n_batch; uint32_t n_pixels; };

#define HK_FLAG_DEPTH(f) ((int)((f) & 0xFFu))
#define HK_FLAG_SPEC 0x100u
#define HK_FLAG_ANYNS 0x200u
#define HK_FLAG_MEDIUM(f) ((f) >> 16)

HK_DEV MatCtx mat_ctx(const DevScene& D) { MatCtx c; c.T = D.T; c.spec_lambdas = D.spec_lambdas; c.spec_values = D.spec_values; c.spec_offsets = D.spec_offsets; c.local = false; return c; }
HK_DEV LightCtx light_ctx(const DevScene& D) {
    LightCtx c; c.T = D.T; c.lights = D.lights; c.n_lights = D.n_lights; c.envmaps = D.envmaps; c.nodes = D.lnodes; c.bit_trails = D.bit_trails;
    c.inf_idx = D.inf_idx; c.n_infinite = D.n_infinite; c.n_bvh = D.n_bvh; c.esc_idx = D.esc_idx; c.n_esc = D.n_esc; return c;
}
HK_DEV MediaCtx media_ctx(const DevScene& D) { MediaCtx c; c.T = D.T; c.media = D.media; c.n_media = D.n_media; c.smem_mask = nullptr; c.smem_medium = 0; return c; }
// the persistent tracking kernels stage the empty-cell mask of DevScene::smem_mask_medium in shared memory (all threads of the block)
HK_DEV MediaCtx media_ctx_staged(const DevScene& D, uint32_t* s_mask) {
    MediaCtx c = media_ctx(D);
    if (D.smem_mask_medium > 0) {
        const uint32_t* __restrict__ g = D.media[D.smem_mask_medium - 1].maj_empty;
        for (uint32_t i = threadIdx.x; i < D.smem_mask_words; i += blockDim.x) s_mask[i] = __ldg(g + i);
        __syncthreads();
        c.smem_mask = s_mask; c.smem_medium = D.smem_mask_medium;
    }
    return c;
}

// Warp-aggregated append to one of several queues: lanes that target the same queue elect a leader, which does a single
// atomicAdd for the group.  Must be called by all 32 lanes (qid < 0 = nothing to push).
HK_DEV void warp_push(uint32_t* counts, uint32_t* const* queues_by_id, int qid, uint32_t value, uint32_t* q_direct = nullptr) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned grp = __match_any_sync(0xFFFFFFFFu, qid);
    if (qid < 0) return;
    unsigned leader = (unsigned)__ffs(grp) - 1u;
    unsigned rank = (unsigned)__popc(grp & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counts + qid, (uint32_t)__popc(grp));
    base = __shfl_sync(grp, base, leader);
    uint32_t* q = q_direct ? q_direct : queues_by_id[qid];
    q[base + rank] = value;
}
// single-queue variant with a predicate (ballot form)
HK_DEV void warp_push1(uint32_t* counter, uint32_t* queue, bool pred, uint32_t value) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if (!pred) return;
    unsigned leader = (unsigned)__ffs(m) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    queue[base + (unsigned)__popc(m & ((1u << lane) - 1u))] = value;
}
// two predicated appends at once: both atomics are in flight before either result is used (one round trip, not two)
HK_DEV void warp_push2(uint32_t* counter0, uint32_t* queue0, bool pred0, uint32_t* counter1, uint32_t* queue1, bool pred1, uint32_t value) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, pred0), m1 = __ballot_sync(0xFFFFFFFFu, pred1);
    const unsigned l0 = m0 ? (unsigned)__ffs(m0) - 1u : 0u, l1 = m1 ? (unsigned)__ffs(m1) - 1u : 0u;
    uint32_t b0 = 0, b1 = 0;
    if (m0 && lane == l0) b0 = atomicAdd(counter0, (uint32_t)__popc(m0));
    if (m1 && lane == l1) b1 = atomicAdd(counter1, (uint32_t)__popc(m1));
    b0 = __shfl_sync(0xFFFFFFFFu, b0, l0); b1 = __shfl_sync(0xFFFFFFFFu, b1, l1);
    if (pred0) queue0[b0 + (unsigned)__popc(m0 & ((1u << lane) - 1u))] = value;
    if (pred1) queue1[b1 + (unsigned)__popc(m1 & ((1u << lane) - 1u))] = value;
}
HK_DEV void count_rays(unsigned long long* ctr, uint32_t mine) {
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(ctr, (unsigned long long)mine);
}

// ---- surface geometry actually consumed downstream (intersection.jl:13-182): pi, flipped geometric normal, shading
// normal.  dpdu/dpdv/dpdus/dpdvs/uv only feed texture filtering, which constant-parameter materials never read. --------
struct Surf { float3 pi, n, ns; float area; uint32_t iface, arealight; };
HK_DEV Surf surface_at(const DevScene& D, uint32_t prim0, float b1, float b2, float3 o, float3 d, float t) {
    Surf s;
    const PrimRef pr = resolve_prim(D, prim0);
    float3 v0, v1, v2;
    prim_vertices(D, pr, v0, v1, v2);
    s.pi = o + d * t;
    float3 cr = cross3(v1 - v0, v2 - v0);
    float3 n = norm3(cr);
    s.area = 0.5f * len3(cr);
    float3 ns = n;
    if (D.normals) {
        const uint32_t i0 = __ldg(D.indices + 3 * (size_t)pr.tri), i1 = __ldg(D.indices + 3 * (size_t)pr.tri + 1), i2 = __ldg(D.indices + 3 * (size_t)pr.tri + 2);
        const float* N = D.normals;
        float3 n0 = f3(__ldg(N + 3 * (size_t)i0), __ldg(N + 3 * (size_t)i0 + 1), __ldg(N + 3 * (size_t)i0 + 2));
        float3 n1 = f3(__ldg(N + 3 * (size_t)i1), __ldg(N + 3 * (size_t)i1 + 1), __ldg(N + 3 * (size_t)i1 + 2));
        float3 n2 = f3(__ldg(N + 3 * (size_t)i2), __ldg(N + 3 * (size_t)i2 + 1), __ldg(N + 3 * (size_t)i2 + 2));
        if (!(isnan(n0.x) || isnan(n1.x) || isnan(n2.x))) {
            if (pr.inst) { n0 = inst_normal(pr.inst->w2o, n0); n1 = inst_normal(pr.inst->w2o, n1); n2 = inst_normal(pr.inst->w2o, n2); }
            float w = 1.0f - b1 - b2;
            ns = norm3(f3(w * n0.x + b1 * n1.x + b2 * n2.x, w * n0.y + b1 * n1.y + b2 * n2.y, w * n0.z + b1 * n1.z + b2 * n2.z));
        }
    }
    s.ns = ns;
    s.n = dot3(n, ns) < 0.0f ? -n : n;
    s.iface = prim_iface(D, pr, prim0);
    s.arealight = prim_arealight(D, pr, prim0);
    return s;
}
// Kd of a textured MatteMaterial at a hit: uv = barycentric interpolation of the vertex uvs (intersection.jl:28-37), bilinear
// texel of the (h, w) column-major image with v flipped and clamped indices (_sample_texture_bilinear, texture-ref.jl:160-190),
// clamp to [0, 1] and uplift (spectral-eval.jl:57-60).  One uncached rgb_to_spectrum per hit.
// uv of a hit = barycentric interpolation of the vertex uvs, (0, 0) without uvs (vp_compute_uv_barycentric, intersection.jl:28-37)
HK_DEV float2 hit_uv(const DevScene& D, uint32_t tri, float b1, float b2) {
    if (!D.uvs) return make_float2(0.0f, 0.0f);
    const uint32_t i0 = __ldg(D.indices + 3 * (size_t)tri), i1 = __ldg(D.indices + 3 * (size_t)tri + 1), i2 = __ldg(D.indices + 3 * (size_t)tri + 2);
    const float2 a = __ldg(reinterpret_cast<const float2*>(D.uvs) + i0), b = __ldg(reinterpret_cast<const float2*>(D.uvs) + i1), c = __ldg(reinterpret_cast<const float2*>(D.uvs) + i2);
    const float w = 1.0f - b1 - b2;
    return make_float2(w * a.x + b1 * b.x + b2 * c.x, w * a.y + b1 * b.y + b2 * c.y);
}
// get_surface_alpha (spectral-eval.jl:3882-3888): alpha of the POINT-sampled Kd texel of a MatteMaterial (_sample_texture_data,
// textures/basic.jl:19-25: idx = trunc(1 + (size - 1) (1 - v, u)), clamped), 1 for every other material (a MixMaterial included)
HK_DEV float surface_alpha(const DevScene& D, uint32_t prim0, float b1, float b2) {
    const PrimRef pr = resolve_prim(D, prim0);
    const HkMaterial& m = D.materials[D.interfaces[prim_iface(D, pr, prim0) - 1].material - 1];
    if (m.type != HK_MAT_MATTE || m.tex[0] <= 0 || (m.flags & HK_MATFLAG_VERTEX_COLORS)) return 1.0f;
    const HkTexture t = D.textures[m.tex[0] - 1];
    if (!t.alpha) return 1.0f;
    const float2 uv = hit_uv(D, pr.tri, b1, b2);
    const int row = clampi((int)(1.0f + (float)(t.h - 1) * (1.0f - uv.y)), 1, t.h), col = clampi((int)(1.0f + (float)(t.w - 1) * uv.x), 1, t.w);
    return __ldg(t.alpha + (size_t)(col - 1) * t.h + (row - 1));
}
// the stochastic alpha test both the trace stage and the shadow rays use (intersection.jl:241-243, 355-358): a PCG32 stream seeded
// from the bits of the ray, so the same ray always gets the same decision
HK_DEV bool alpha_skips(const DevScene& D, uint32_t prim0, float b1, float b2, float3 o, float3 d) {
    const float alpha = surface_alpha(D, prim0, b1, b2);
    if (!(alpha < 1.0f)) return false;
    Pcg32 rng = pcg32_init(hash_f3(o), hash_f3(d));
    return pcg32_f32(rng) > alpha;
}
// _sample_texture_bilinear (texture-ref.jl:160-190) on an (h, w) column-major RGB image: v flipped, clamped indices, no wrap, no clamp of the value
HK_DEV void tex_bilinear(const HkTexture& t, float2 uv, float* rgb) {
    const float px = uv.x * (float)(t.w - 1) + 1.0f, py = (1.0f - uv.y) * (float)(t.h - 1) + 1.0f;
    const float flx = floorf(px), fly = floorf(py);
    const int x0 = clampi(floor_i(px), 1, t.w), x1 = clampi(floor_i(px) + 1, 1, t.w), y0 = clampi(floor_i(py), 1, t.h), y1 = clampi(floor_i(py) + 1, 1, t.h);
    const float fx = px - flx, fy = py - fly;
    const float* T = t.rgb;
    const size_t o00 = 3 * ((size_t)(x0 - 1) * t.h + (y0 - 1)), o10 = 3 * ((size_t)(x1 - 1) * t.h + (y0 - 1));
    const size_t o01 = 3 * ((size_t)(x0 - 1) * t.h + (y1 - 1)), o11 = 3 * ((size_t)(x1 - 1) * t.h + (y1 - 1));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float c0 = __ldg(T + o00 + k) * (1.0f - fx) + __ldg(T + o10 + k) * fx, c1 = __ldg(T + o01 + k) * (1.0f - fx) + __ldg(T + o11 + k) * fx;
        rgb[k] = c0 * (1.0f - fy) + c1 * fy;
    }
}
HK_DEV Spec textured_kd(const DevScene& D, const HkMaterial& m, uint32_t prim0, float b1, float b2, float4 lam) {
    const PrimRef pr = resolve_prim(D, prim0);
    const HkTexture t = D.textures[m.tex[0] - 1];
    if (m.flags & HK_MATFLAG_VERTEX_COLORS) {
        // eval_tex(::VertexColorTexture, tfc), texture-ref.jl:240-245: data[1, fi] b1 + data[2, fi] b2 + data[3, fi] b3 over a (3, n_faces)
        // table of per-face corner colours; fi = TriangleMeta.primitive_index (the face within its mesh)
        const uint32_t face = pr.inst ? prim0 - pr.inst->prim_base + 1u : __ldg(D.tri_meta + 3 * (size_t)prim0 + 1);
        const float* T = t.rgb + 9 * (size_t)(face - 1u);
        const float w0 = 1.0f - b1 - b2;
        float rgb[3];
#pragma unroll
        for (int k = 0; k < 3; k++) rgb[k] = clampf(__ldg(T + k) * w0 + __ldg(T + 3 + k) * b1 + __ldg(T + 6 + k) * b2, 0.0f, 1.0f);
        return pre_bounded(make_pre_bounded(D.T, rgb[0], rgb[1], rgb[2]), lam);
    }
    float rgb[3];
    tex_bilinear(t, hit_uv(D, pr.tri, b1, b2), rgb);
    return pre_bounded(make_pre_bounded(D.T, clampf(rgb[0], 0.0f, 1.0f), clampf(rgb[1], 0.0f, 1.0f), clampf(rgb[2], 0.0f, 1.0f)), lam);
}
// Textured parameters of any material (eval_tex(textures, mat.<param>, tfc) at every parameter read of spectral-eval.jl): a per-hit
// copy of the material whose textured RGB / scalar parameters hold the texel values at the hit; everything downstream -- clamps,
// `albedo == 0` tests, uplifts -- then treats them exactly like the constants they replace.  Out of line: constant-parameter
// materials never get here.
static __device__ __noinline__ void resolve_material_textures(const DevScene& D, const HkMaterial& g, uint32_t prim0, float b1, float b2, HkMaterial& out) {
    out = g;
    const PrimRef pr = resolve_prim(D, prim0);
    const float2 uv = hit_uv(D, pr.tri, b1, b2);
    float rgb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (g.tex[k] <= 0 || (k == 0 && g.type == HK_MAT_MATTE)) continue;      // (Matte.Kd: textured_kd, its own shading class)
        tex_bilinear(D.textures[g.tex[k] - 1], uv, rgb);
        float* dst = k == 0 ? out.rgb0 : (k == 1 ? out.rgb1 : out.rgb2);
        dst[0] = rgb[0]; dst[1] = rgb[1]; dst[2] = rgb[2];
    }
    for (int k = 0; k < 8; k++) {
        if (g.ftex[k] <= 0) continue;
        tex_bilinear(D.textures[g.ftex[k] - 1], uv, rgb);
        out.f[k] = rgb[0];
    }
}
HK_DEV bool material_has_textures(const HkMaterial& m) {
    return ((m.tex[0] > 0 && m.type != HK_MAT_MATTE) | (m.tex[1] > 0) | (m.tex[2] > 0) | ((m.ftex[0] | m.ftex[1] | m.ftex[2] | m.ftex[3] | m.ftex[4] | m.ftex[5] | m.ftex[6] | m.ftex[7]) > 0)) != 0;
}
HK_DEV float3 geometric_normal(const DevScene& D, uint32_t prim0) {
    float3 v0, v1, v2;
    prim_vertices(D, resolve_prim(D, prim0), v0, v1, v2);
    return norm3(cross3(v1 - v0, v2 - v0));
}
HK_DEV int material_type_of_prim(const DevScene& D, uint32_t prim0) {
    uint32_t mi = prim_iface(D, prim0);
    uint32_t mat = __ldg(&D.interfaces[mi - 1].material);
    return __ldg(&D.materials[mat - 1].type);
}

// =====================================================================================================================
// stage kernels
// =====================================================================================================================
HK_DEV int slot_sample_idx(const PassArgs& A, uint32_t slot) { return A.first_sample + A.stride * (int)(slot / A.n_pixels); }

#ifdef HK_TU_CORE
// vp_generate_camera_rays_kernel!, volpath.jl:125-205.  One thread per slot; ray queue 0 becomes the identity.
__global__ void __launch_bounds__(256) k_camera(const __grid_constant__ DevScene D, PathState S, PassArgs A, const uint32_t* __restrict__ camera_medium_dev) {
    const uint32_t camera_medium = __ldg(camera_medium_dev);      // detect_camera_medium's result, left on the device
    const uint32_t n_slots = A.n_pixels * (uint32_t)A.n_batch;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
        const uint32_t pix = slot % A.n_pixels;
        const int sample_idx = slot_sample_idx(A, slot);
        const int x = (int)(pix % (uint32_t)D.width) + 1, y = (int)(pix / (uint32_t)D.width) + 1;
        float wu = zsobol_1d(D.sobol, x, y, sample_idx, 1, HK_SOBOL_SLOT_CAMERA(0), pix);
        float2 jit = zsobol_2d(D.sobol, x, y, sample_idx, 3, HK_SOBOL_SLOT_CAMERA(1), pix);
        // the time sample (dim 4) is drawn by the reference but VolPath's ray drops it (volpath.jl:184)
        float2 lens = D.camera.lens_radius > 0.0f ? zsobol_2d(D.sobol, x, y, sample_idx, 6, HK_SOBOL_SLOT_CAMERA(2), pix) : make_float2(0.0f, 0.0f);
        float fx, fy, fw;
        filter_sample(D.filter, jit, fx, fy, fw);
        float4 lam, pdf;
        sample_wavelengths_visible(wu, lam, pdf);
        float3 o, d;
        camera_generate_ray(D.camera, (float)x + 0.5f + fx, (float)D.height - (float)y + 1.0f + 0.5f + fy, lens, o, d);
        S.fweight[slot] = fw; S.lambda[slot] = lam; S.lpdf[slot] = pdf;
        S.ray_a[slot] = make_float4(o.x, o.y, o.z, d.x);
        S.ray_b[slot] = make_float4(d.y, d.z, HK_INF, 0.0f);
        S.beta[slot] = sp(1.0f); S.r_u[slot] = sp(1.0f); S.r_l[slot] = sp(1.0f); S.L[slot] = sp(0.0f);
        S.flags[slot] = camera_medium << 16;
        S.q_ray[0][slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { S.counts[HK_C_RAY0] = n_slots; S.counts[HK_C_RAY1] = 0; }
}

// Uplift cache (DevTables::mat_pre / light_pre / med_pre): one thread per material / light / medium, run at upload.
// T arrives with null cache pointers, so the make_* functions evaluate directly.
__global__ void __launch_bounds__(128) k_precompute_uplifts(DevTables T, const HkMaterial* __restrict__ mats, uint32_t n_mats, float4* __restrict__ mat_pre,
                                                             const HkLight* __restrict__ lights, uint32_t n_lights, float4* __restrict__ light_pre,
                                                             const DevMedium* __restrict__ media, uint32_t n_media, float4* __restrict__ med_pre) {
    const uint32_t n = n_mats + n_lights + n_media;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < n_mats) {
            const HkMaterial& m = mats[i];
            const bool rgb = (m.type >= 1 && m.type < HK_MAX_MAT_TYPES && m.type != HK_MAT_THIN_DIELECTRIC) || m.type == HK_MAT_COATED_CONDUCTOR || m.type == HK_MAT_COATED_DIFFUSE_TRANSMISSION;
            mat_pre[2 * i] = rgb ? mat_pre_compute(T, m, 0) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            mat_pre[2 * i + 1] = rgb ? mat_pre_compute(T, m, 1) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else if (i < n_mats + n_lights) {
            const uint32_t k = i - n_mats;
            const HkLight& L = lights[k];
            light_pre[2 * k] = make_pre_illuminant(T, L.rgb[0], L.rgb[1], L.rgb[2]);
            light_pre[2 * k + 1] = make_pre_bounded(T, L.rgb[0] * L.scale, L.rgb[1] * L.scale, L.rgb[2] * L.scale);
        } else {
            const uint32_t k = i - n_mats - n_lights;
            const DevMedium& M = media[k];
            med_pre[3 * k] = make_pre_unbounded(T, M.sigma_a[0], M.sigma_a[1], M.sigma_a[2]);
            med_pre[3 * k + 1] = make_pre_unbounded(T, M.sigma_s[0], M.sigma_s[1], M.sigma_s[2]);
            med_pre[3 * k + 2] = make_pre_unbounded(T, M.Le[0], M.Le[1], M.Le[2]);
        }
    }
}

// empty-cell mask of a majorant grid (DevMedium::maj_empty): one thread per 32 cells
__global__ void __launch_bounds__(256) k_majorant_mask(const float* __restrict__ grid, uint32_t n_cells, uint32_t* __restrict__ mask) {
    const uint32_t n_words = (n_cells + 31u) / 32u;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32u; b++) { const uint32_t c = 32u * w + b; if (c < n_cells && grid[c] == 0.0f) bits |= 1u << b; }
        mask[w] = bits;
    }
}
// Majorant grid of a grid / RGB grid / NanoVDB medium, built on the device from the uploaded voxels (HkMedium.majorant == NULL):
//   GridMedium     build_majorant_grid          media.jl:1459-1496   max density over the voxels cell [i, i+1)/res maps to
//   RGBGridMedium  build_rgb_majorant_grid      media.jl:1123-1183   sigma_scale * (max MaxValue(sigma_a) + max MaxValue(sigma_s)); absent grid = 1
//   NanoVDBMedium  build_nanovdb_majorant_grid  nanovdb.jl:1174-1235 max tree value over the index box of the cell's two corners
//                                                                    +- 1 voxel, clipped to [index_min, index_max]
// One block per majorant cell; the threads stride over the cell's voxel box and the block reduces with Julia's NaN-propagating max
// (a max is order independent, so the grid equals the serial loop's bit for bit).  The index ranges are the reference's: integer
// floor / ceil of i*n/res (exact in integers), and for NanoVDB the cell corners in f32 with every product and sum rounded on its own.
struct MajBuild { int32_t idx_min[3], idx_max[3]; float bmin[3], bmax[3]; };
HK_DEV void maj_range(int i, int n, int r, int& a, int& b) {      // 0-based [a, b): max(1, floor(i n / r) + 1) .. min(n, ceil((i+1) n / r))
    a = (int)(((long long)i * n) / r);
    b = (int)((((long long)(i + 1)) * n + r - 1) / r); if (b > n) b = n;
}
__global__ void __launch_bounds__(128) k_build_majorant(DevMedium M, MajBuild P, float* __restrict__ out) {
    __shared__ float red[2][4];
    const int rx = M.mres[0], ry = M.mres[1];
    const int cell = blockIdx.x, ix = cell % rx, iy = (cell / rx) % ry, iz = cell / (rx * ry);
    float m0 = 0.0f, m1 = 0.0f;
    if (M.type == HK_MEDIUM_NANOVDB) {
        const int ci[3] = {ix, iy, iz};
        float p0[3], p1[3], q0[3], q1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float diag = __fsub_rn(P.bmax[k], P.bmin[k]), r = (float)M.mres[k];
            p0[k] = __fsub_rn(__fadd_rn(P.bmin[k], __fdiv_rn(__fmul_rn(diag, (float)ci[k]), r)), M.vec[k]);
            p1[k] = __fsub_rn(__fadd_rn(P.bmin[k], __fdiv_rn(__fmul_rn(diag, (float)(ci[k] + 1)), r)), M.vec[k]);
        }
        int lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {      // world_to_index_f_raw, nanovdb.jl:1238-1246: (m1 px + m2 py) + m3 pz
            q0[k] = __fadd_rn(__fadd_rn(__fmul_rn(M.inv_mat[3 * k], p0[0]), __fmul_rn(M.inv_mat[3 * k + 1], p0[1])), __fmul_rn(M.inv_mat[3 * k + 2], p0[2]));
            q1[k] = __fadd_rn(__fadd_rn(__fmul_rn(M.inv_mat[3 * k], p1[0]), __fmul_rn(M.inv_mat[3 * k + 1], p1[1])), __fmul_rn(M.inv_mat[3 * k + 2], p1[2]));
            lo[k] = max((int)floorf(__fsub_rn(fminf(q0[k], q1[k]), 1.0f)), P.idx_min[k]);
            hi[k] = min((int)ceilf(__fadd_rn(fmaxf(q0[k], q1[k]), 1.0f)), P.idx_max[k]);
        }
        if (lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) {
            const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
            LeafCache lc; lc.valid = false;
            for (int v = threadIdx.x; v < ex * ey * ez; v += blockDim.x) {      // z fastest: neighbouring threads share leaves
                const int z = v % ez, y = (v / ez) % ey, x = v / (ez * ey);
                m0 = jl_max(m0, nvdb_value(M, lc, lo[0] + x, lo[1] + y, lo[2] + z));
            }
        }
    } else {
        int x0, x1, y0, y1, z0, z1;
        maj_range(ix, M.dres[0], M.mres[0], x0, x1); maj_range(iy, M.dres[1], M.mres[1], y0, y1); maj_range(iz, M.dres[2], M.mres[2], z0, z1);
        const int ex = max(x1 - x0, 0), ey = max(y1 - y0, 0), ez = max(z1 - z0, 0);
        for (int v = threadIdx.x; v < ex * ey * ez; v += blockDim.x) {
            const int x = x0 + v % ex, y = y0 + (v / ex) % ey, z = z0 + v / (ex * ey);
            const size_t at = (size_t)x + (size_t)M.dres[0] * ((size_t)y + (size_t)M.dres[1] * (size_t)z);
            if (M.type == HK_MEDIUM_GRID) m0 = jl_max(m0, __ldg(M.density + at));
            else {
                if (M.rgb_a) m0 = jl_max(m0, jl_max(jl_max(__ldg(M.rgb_a + 3 * at), __ldg(M.rgb_a + 3 * at + 1)), __ldg(M.rgb_a + 3 * at + 2)));
                if (M.rgb_s) m1 = jl_max(m1, jl_max(jl_max(__ldg(M.rgb_s + 3 * at), __ldg(M.rgb_s + 3 * at + 1)), __ldg(M.rgb_s + 3 * at + 2)));
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { m0 = jl_max(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, o)); m1 = jl_max(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, o)); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = m0; red[1][threadIdx.x >> 5] = m1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; w++) { m0 = jl_max(m0, red[0][w]); m1 = jl_max(m1, red[1][w]); }
        if (M.type == HK_MEDIUM_RGBGRID) { if (!M.rgb_a) m0 = 1.0f; if (!M.rgb_s) m1 = 1.0f; m0 = __fmul_rn(M.sigma_scale, __fadd_rn(m0, m1)); }
        out[cell] = m0;
    }
}
// dense mirror of a NanoVDB tree (DevMedium::dense): one thread per voxel of the index box, the value the tree walk returns
__global__ void __launch_bounds__(256) k_nvdb_densify(DevMedium M, float* __restrict__ out) {
    const size_t n = (size_t)M.dn_ext[0] * M.dn_ext[1] * M.dn_ext[2];
    LeafCache lc; lc.valid = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % (size_t)M.dn_ext[2]), y = (int)((i / (size_t)M.dn_ext[2]) % (size_t)M.dn_ext[1]), x = (int)(i / ((size_t)M.dn_ext[2] * M.dn_ext[1]));
        out[i] = nvdb_value(M, lc, M.dn_min[0] + x, M.dn_min[1] + y, M.dn_min[2] + z);
    }
}
// light-BVH nodes in their device form (DevLNode, hk_lights.cuh): the point-independent part of node_importance, once per upload
__global__ void __launch_bounds__(256) k_prepare_lnodes(const HkLightBVHNode* __restrict__ in, uint32_t n, DevLNode* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = prepare_lnode(in[i]);
}
// ZSobol prefix cache (SobolParams::top, hk_math.cuh): one pass per (resolution, seed), not per sample.
// dims[slot] = the sampler dimension of cache slot `slot`.
__global__ void __launch_bounds__(256) k_sobol_prefix(uint32_t* __restrict__ top, uint4* __restrict__ dimhash, const int32_t* __restrict__ dims, int32_t n_slots,
                                                       uint32_t n_pixels, int32_t width, int32_t log2_spp, int32_t nb4, uint32_t seed) {
    const size_t total = (size_t)n_slots * n_pixels;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t)(i / n_pixels), pix = (uint32_t)(i % n_pixels);
        const int32_t dim = __ldg(dims + slot);
        top[i] = zsobol_prefix(pix % (uint32_t)width + 1u, pix / (uint32_t)width + 1u, dim, log2_spp, nb4);
        if (pix == 0) {
            const uint64_t b2 = hash_dim_seed(dim + 2, seed);
            dimhash[slot] = make_uint4((uint32_t)hash_dim_seed(dim + 1, seed), (uint32_t)b2, (uint32_t)(b2 >> 32), 0u);
        }
    }
}

// reset_iteration_queues!, volpath-state.jl:214-222 (+ the next ray queue and the traversal cursors)
// keep_par >= 0: the shadow-pass counters of that parity belong to the previous bounce's shadow kernel, which may still be
// running on its own stream: leave them alone
__global__ void k_reset_bounce(PathState S, int cur, int keep_par) {      // launched with HK_N_COUNTERS threads
    int i = threadIdx.x;
    if (keep_par >= 0 && (i == HK_CI_SHADOW(keep_par) || i == HK_CI_CURSOR_SHADOW(keep_par) || i == HK_CI_TOTAL_HITS(keep_par))) return;
    if (i < HK_N_COUNTERS && i != (HK_C_RAY0 + cur)) S.counts[i] = 0;
}
#endif  // HK_TU_CORE
// append from whichever lanes are here (a divergent region of a persistent loop): lanes that arrive together share one atomic
HK_DEV void push_active(uint32_t* counter, uint32_t* queue, uint32_t value) {
    const unsigned m = __activemask(), lane = threadIdx.x & 31u;
    const unsigned leader = (unsigned)__ffs(m) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    queue[base + (unsigned)__popc(m & ((1u << lane) - 1u))] = value;
}
// claim `popc(idle)` consecutive work items for the idle lanes of a warp; returns this lane's index (only meaningful for idle lanes)
HK_DEV uint32_t claim_for_idle(uint32_t* cursor, unsigned idle) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned leader = (unsigned)__ffs(idle) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
    base = __shfl_sync(idle, base, leader);
    return base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
}

HK_DEV uint32_t* queue_of(const PathState& S, int qid) {
    if (qid >= HK_C_ALPHA_N0) return S.q_alpha[(qid - HK_C_ALPHA_N0) & 1];
    if (qid == HK_C_MEDIUM) return S.q_medium;
    if (qid == HK_C_ESCAPED) return S.q_escaped;
    if (qid >= HK_C_HIT1) return S.q_hit[8 + qid - HK_C_HIT1];
    if (qid >= HK_C_HIT0) return S.q_hit[qid - HK_C_HIT0];
    return nullptr;
}

// vp_trace_rays_kernel!, intersection.jl:188-269, split in two: k_trace = closest hit for every queued ray
// (persistent, per-lane refill), k_route = classification into the medium / escaped / per-material queues.
struct QueueRayIO {
    const PathState& S; const uint32_t* __restrict__ q;
    HK_DEV uint32_t load(uint32_t idx, float3& o, float3& d, float& tm) const {
        const uint32_t slot = q[idx];
        const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
        o = f3(ra.x, ra.y, ra.z); d = f3(ra.w, rb.x, rb.y); tm = rb.z;
        return slot;
    }
    HK_DEV void store(uint32_t slot, const HitRec& h) const { S.hit[slot] = make_float4(h.t, __uint_as_float(h.prim1), h.b1, h.b2); }
};
#ifdef HK_TU_TRACE
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(HK_TRACE_THREADS, INST ? HK_TRACE_BLOCKS_PER_SM_INST : HK_TRACE_BLOCKS_PER_SM) k_trace(const __grid_constant__ DevScene D, PathState S, int cur, int round, unsigned long long* work) {
    HK_TRACE_SMEM(INST);
    // round > 0: a retrace round of the alpha loop (rays whose previous hit was skipped, restarted just behind that surface)
    const uint32_t n = S.counts[round == 0 ? HK_C_RAY0 + cur : HK_C_ALPHA_N0 + round];
    uint32_t traced = 0, wn = 0, wt = 0;
    QueueRayIO io{S, round == 0 ? S.q_ray[cur] : S.q_alpha[round & 1]};
    trace_queue<false, COUNT, INST>(D.bvh, sm_stack + threadIdx.x, sm_wray + threadIdx.x, n, S.counts + (round == 0 ? HK_C_CURSOR_TRACE : HK_C_ALPHA_CUR0 + round), io, traced, wn, wt);
    count_rays(S.rays_traced, traced);
    if (COUNT) { count_rays(work, traced); count_rays(work + 1, wn); count_rays(work + 2, wt); }
}
#endif  // HK_TU_TRACE
// in-medium rays go to delta tracking with their hit record; the vacuum alpha loop (:224-266) ends on its first iteration
// because every constant-parameter material has alpha == 1 (spectral-eval.jl:3882-3888)
// ---- MixMaterial, src/materials/mix-material.jl: mix_hash_float :114-158 (the UInt32 shifts truncate, the SetKey shifts
// are 64-bit), choose_material :178-196, resolve_mix_material :253-268 (<= 8 levels) ---------------
HK_DEV float mix_hash_float(float3 p, float3 wo, uint32_t type1, uint32_t vec1, uint32_t type2, uint32_t vec2) {
    uint64_t h = 0;
    h ^= (uint64_t)__float_as_uint(p.x);
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)(uint32_t)(__float_as_uint(p.y) << 4);
    h *= 0x1b873593ull;
    h ^= (uint64_t)(uint32_t)(__float_as_uint(p.z) << 8);
    h ^= (uint64_t)(uint32_t)(__float_as_uint(wo.x) << 16);
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)__float_as_uint(wo.y);
    h *= 0x1b873593ull;
    h ^= (uint64_t)(uint32_t)(__float_as_uint(wo.z) << 12);
    h ^= (uint64_t)type1 << 24;
    h ^= (uint64_t)vec1;
    h *= 0xcc9e2d51ull;
    h ^= (uint64_t)type2 << 28;
    h ^= (uint64_t)vec2 << 4;
    h *= 0x1b873593ull;
    h = mix_bits(h);
    return (float)(uint32_t)(h & 0xFFFFFFFFull) * 2.3283064365386963e-10f;
}
// amount: the constant, or eval_tex(ctx, mix.amount, uv) = the bilinear texel at the hit's uv (choose_material :183)
HK_DEV uint32_t resolve_mix_material(const DevScene& D, uint32_t idx, float3 p, float3 wo, uint32_t prim0, float b1, float b2) {
    const HkMaterial* __restrict__ materials = D.materials;
    bool have_uv = false; float2 uv = make_float2(0.0f, 0.0f);
    for (int it = 0; it < 8; it++) {
        const HkMaterial& m = materials[idx - 1];
        if (m.type != HK_MAT_MIX) return idx;
        float amt = m.f[0];
        if (m.ftex[0] > 0) {
            if (!have_uv) { uv = hit_uv(D, resolve_prim(D, prim0).tri, b1, b2); have_uv = true; }
            float rgb[3]; tex_bilinear(D.textures[m.ftex[0] - 1], uv, rgb); amt = rgb[0];
        }
        if (amt <= 0.0f) idx = (uint32_t)m.ival[0];
        else if (amt >= 1.0f) idx = (uint32_t)m.ival[1];
        else {
            const float u = mix_hash_float(p, wo, m.flags & 0xFFu, (uint32_t)m.spec[0], (m.flags >> 8) & 0xFFu, (uint32_t)m.spec[1]);
            idx = amt < u ? (uint32_t)m.ival[0] : (uint32_t)m.ival[1];
        }
    }
    return idx;
}
// queue of a surface hit: the material type rides in the hit record; a MixMaterial is resolved here, at intersection time,
// from the hit point and the outgoing direction (surface-eval.jl:162-168), and the chosen material is left for k_shade
HK_DEV int hit_queue_id(const DevScene& D, const PathState& S, uint32_t slot, uint32_t hit_bits, float t_hit) {
    uint32_t mtype = HK_HIT_MTYPE(hit_bits);
    if (mtype == HK_MAT_MIX) {
        const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
        const float3 o = f3(ra.x, ra.y, ra.z), d = f3(ra.w, rb.x, rb.y);
        const uint32_t prim0 = HK_HIT_PRIM1(hit_bits) - 1u;
        const uint32_t mi = prim_iface(D, prim0);
        const float4 hr = S.hit[slot];      // (t, bits, b1, b2): the barycentrics are only needed for a textured amount
        const uint32_t res = resolve_mix_material(D, D.interfaces[mi - 1].material, o + d * t_hit, -d, prim0, hr.z, hr.w);
        S.res_mat[slot] = res;
        mtype = shade_class(D.materials[res - 1]);
        if (mtype == HK_MAT_MIX) return -1;            // a mix chain deeper than 8 levels: the reference would shade a MixMaterial (undefined); dropped
    }
    return HK_HIT_COUNTER((int)HK_TYPE_QUEUE(mtype));
}

// the alpha test of a routed hit; on a skip the slot's ray restarts behind the surface (same direction, t_max = Inf)
HK_DEV bool route_alpha_skip(const DevScene& D, const PathState& S, uint32_t slot, uint32_t hit_bits, float t_hit) {
    const float4 hr = S.hit[slot];
    const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
    const float3 o = f3(ra.x, ra.y, ra.z), d = f3(ra.w, rb.x, rb.y);
    const uint32_t prim0 = HK_HIT_PRIM1(hit_bits) - 1u;
    if (!alpha_skips(D, prim0, hr.z, hr.w, o, d)) return false;
    const float3 pi = o + d * t_hit;
    const float3 n = geometric_normal(D, prim0);
    const float3 no = pi + (dot3(d, n) > 0.0f ? n : -n) * 1.0e-4f;
    S.ray_a[slot] = make_float4(no.x, no.y, no.z, d.x);
    S.ray_b[slot] = make_float4(d.y, d.z, HK_INF, 0.0f);
    return true;
}
#ifdef HK_TU_CORE
// writes the shading class (material type; 11 = textured matte) of every BVH triangle into the spare word of its record (HitRec)
__global__ void __launch_bounds__(256) k_patch_tri_types(float4* __restrict__ tris, uint32_t n_tris, const uint32_t* __restrict__ tri_meta,
                                                          const HkMediumInterface* __restrict__ interfaces, const HkMaterial* __restrict__ materials) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tris; i += gridDim.x * blockDim.x) {
        const uint32_t prim0 = __float_as_uint(tris[3 * (size_t)i].w);
        const uint32_t mi = tri_meta[3 * (size_t)prim0];
        const uint32_t type = shade_class(materials[interfaces[mi - 1].material - 1]);
        tris[3 * (size_t)i + 1].w = __uint_as_float(type & 0xFu);
    }
}
// instanced scenes: the shading class rides in the instance's leaf record (word 2 of its fourth float4, already shifted to bits 28-31)
__global__ void __launch_bounds__(256) k_patch_inst_types(float4* __restrict__ inst_recs, uint32_t n_inst, const DevInstance* __restrict__ instances,
                                                           const HkMediumInterface* __restrict__ interfaces, const HkMaterial* __restrict__ materials) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += gridDim.x * blockDim.x) {
        const uint32_t orig = __float_as_uint(inst_recs[4 * (size_t)i + 3].w);
        const uint32_t type = shade_class(materials[interfaces[instances[orig].iface - 1].material - 1]);
        inst_recs[4 * (size_t)i + 3].z = __uint_as_float((type & 0xFu) << 28);
    }
}
// Appends are aggregated per BLOCK: every thread classifies HK_ROUTE_PER_THREAD rays, takes its positions from shared-memory
// counters, and one thread per queue then reserves the block's range with a single global atomicAdd (1 per queue per 1024
// rays).  With warp-level aggregation alone the ~0.5 M same-address atomics per sample were what the kernel waited for.
#define HK_ROUTE_PER_THREAD 4
// round: 0 = the rays of the bounce; r > 0 = the rays re-traced in alpha round r.  A hit on an alpha-tested surface that the
// stochastic test skips (vp_trace_rays_kernel!, intersection.jl:221-266) restarts its ray 1e-4 behind the surface and queues it
// for round r + 1 -- no depth consumed; a ray still being skipped in the 16th round is absorbed.
__global__ void __launch_bounds__(256) k_route(const __grid_constant__ DevScene D, PathState S, int cur, int par, int round) {
    __shared__ uint32_t s_cnt[HK_N_COUNTERS], s_base[HK_N_COUNTERS];
    const uint32_t n = S.counts[round == 0 ? HK_C_RAY0 + cur : HK_C_ALPHA_N0 + round];
    const uint32_t chunk = 256u * HK_ROUTE_PER_THREAD;
    const uint32_t* __restrict__ q = round == 0 ? S.q_ray[cur] : S.q_alpha[round & 1];
    for (uint32_t c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {
        if (threadIdx.x < HK_N_COUNTERS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int qid[HK_ROUTE_PER_THREAD]; uint32_t slot[HK_ROUTE_PER_THREAD], pos[HK_ROUTE_PER_THREAD];
#pragma unroll
        for (int r = 0; r < HK_ROUTE_PER_THREAD; r++) {
            const uint32_t i = c0 + r * 256u + threadIdx.x;
            qid[r] = -1; slot[r] = 0; pos[r] = 0;
            if (i < n) {
                slot[r] = q[i];
                const float2 hty = *reinterpret_cast<const float2*>(&S.hit[slot[r]]);
                const uint32_t hb = __float_as_uint(hty.y);
                if (D.n_media > 0 && HK_FLAG_MEDIUM(S.flags[slot[r]]) != 0) qid[r] = HK_C_MEDIUM;
                else if (HK_HIT_PRIM1(hb) == 0) qid[r] = HK_C_ESCAPED;
                else if (D.has_alpha && HK_HIT_MTYPE(hb) == HK_SHADE_MATTE_TEX && route_alpha_skip(D, S, slot[r], hb, hty.x)) qid[r] = round + 1 < HK_ALPHA_ROUNDS ? HK_C_ALPHA_N0 + round + 1 : -1;
                else qid[r] = hit_queue_id(D, S, slot[r], hb, hty.x);
            }
        }
#pragma unroll
        for (int r = 0; r < HK_ROUTE_PER_THREAD; r++) if (qid[r] >= 0) pos[r] = atomicAdd(&s_cnt[qid[r]], 1u);
        __syncthreads();
        if (threadIdx.x < HK_N_COUNTERS) {
            const uint32_t c = s_cnt[threadIdx.x];
            if (c) s_base[threadIdx.x] = atomicAdd(S.counts + threadIdx.x, c);
        }
        if (threadIdx.x == 32) {   // total surface hits of the bounce (the reference's `n_hits > 0` shadow-pass condition)
            uint32_t h = 0;
            for (int t = 0; t < HK_N_HIT_QUEUES; t++) h += s_cnt[HK_HIT_COUNTER(t)];
            if (h) { atomicAdd(S.counts + HK_CI_TOTAL_HITS(par), h); atomicAdd(S.path_vertices, (unsigned long long)h); }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < HK_ROUTE_PER_THREAD; r++) if (qid[r] >= 0) queue_of(S, qid[r])[s_base[qid[r]] + pos[r]] = slot[r];
        __syncthreads();
    }
}

// Ray-queue ordering for traversal coherence: continuation rays leave the shading kernels in hit-queue order -- neighbouring origins,
// unrelated directions.  Each block regroups its tile of HK_SORT_TILE queue entries by direction octant, in place (a counting sort in
// shared memory): the lanes of a traversal warp then walk the same children in the same order far more often.  Queue order has no
// effect on any result (every slot is independent; the film is summed per slot in sample order).
#ifndef HK_SORT_TILE
#define HK_SORT_TILE 2048
#endif
__global__ void __launch_bounds__(256) k_sort_rays(PathState S, int cur) {
    __shared__ uint32_t s_slot[HK_SORT_TILE];
    __shared__ uint8_t s_oct[HK_SORT_TILE];
    __shared__ uint32_t s_cnt[8], s_off[8];
    const uint32_t n = S.counts[HK_C_RAY0 + cur];
    uint32_t* __restrict__ q = S.q_ray[cur];
    for (uint32_t t0 = blockIdx.x * HK_SORT_TILE; t0 < n; t0 += gridDim.x * HK_SORT_TILE) {
        if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t pos[HK_SORT_TILE / 256];
#pragma unroll
        for (int k = 0; k < HK_SORT_TILE / 256; k++) {
            const uint32_t j = k * 256u + threadIdx.x, i = t0 + j;
            pos[k] = 0;
            if (i < n) {
                const uint32_t slot = q[i];
                const float dx = S.ray_a[slot].w; const float2 dyz = *reinterpret_cast<const float2*>(&S.ray_b[slot]);
                const uint32_t oct = (dx >= 0.0f ? 4u : 0u) | (dyz.x >= 0.0f ? 2u : 0u) | (dyz.y >= 0.0f ? 1u : 0u);
                s_slot[j] = slot; s_oct[j] = (uint8_t)oct;
                pos[k] = atomicAdd(&s_cnt[oct], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t a = 0; for (int o = 0; o < 8; o++) { s_off[o] = a; a += s_cnt[o]; } }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < HK_SORT_TILE / 256; k++) {
            const uint32_t j = k * 256u + threadIdx.x, i = t0 + j;
            if (i < n) q[t0 + s_off[s_oct[j]] + pos[k]] = s_slot[j];
        }
        __syncthreads();
    }
}

// vp_handle_escaped_rays_kernel!, intersection.jl:622-668
__global__ void __launch_bounds__(256) k_escaped(const __grid_constant__ DevScene D, PathState S) {
    const uint32_t n = S.counts[HK_C_ESCAPED];
    if (D.n_lights <= 0) return;
    LightCtx LC = light_ctx(D);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = S.q_escaped[i];
        float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
        float3 d = f3(ra.w, rb.x, rb.y);
        float4 lam = S.lambda[slot];
        Spec contrib = S.beta[slot] * escaped_Le(LC, d, lam);
        if (sp_black(contrib)) continue;
        const uint32_t fl = S.flags[slot];
        Spec ru = S.r_u[slot], fin;
        if (HK_FLAG_DEPTH(fl) == 0 || (fl & HK_FLAG_SPEC)) fin = contrib / sp_avg(ru);
        else {
            float lcp = 1.0f / (float)D.n_lights;
            Spec rl = S.r_l[slot] * lcp * env_light_pdf(LC, d);
            float den = sp_avg(ru + rl);
            fin = den > 1.0e-10f ? contrib / den : contrib / sp_avg(ru);
        }
        S.L[slot] = S.L[slot] + fin;
    }
}

#endif  // HK_TU_CORE
// russian_roulette_spectral, material-dispatch.jl:263-287
HK_DEV bool russian_roulette(Spec& beta, int depth, float rr) {
    if (depth <= 3) return true;
    float q = fmaxf(0.05f, 1.0f - sp_maxc(beta));
    if (rr < q) return false;
    beta = beta * (1.0f / (1.0f - q));
    return true;
}

#ifdef HK_TU_LIGHTS
// The light half of surface shading, for the hits of ALL material queues of a bounce in one kernel: emissive-hit MIS
// (HandleEmissiveIntersection, surface-eval.jl:147-220) and the light sample of next-event estimation -- BVH light selection
// (bvh-light-sampler.jl:105-170) + sample_light (lights.jl:39-290) -- whose result k_shade<TYPE> picks up from the per-slot record
// nee_a / nee_b / nee_c.  It used to be inlined into every k_shade<TYPE>: ~4 000 of the ~7 000 SASS instructions of each shading
// kernel (110-230 KB against a 32 KB L1.5 instruction cache; ncu: stall_no_instruction 11 warps per issue on C3's conductor kernel,
// issue slots 20 % busy).  Same functions on the same inputs in the same order: same bits.
__global__ void __launch_bounds__(128, 4) k_hit_lights(const __grid_constant__ DevScene D, PathState S, PassArgs A) {
    LightCtx LC = light_ctx(D);
    for (int q = 0; q < HK_N_HIT_QUEUES; q++) {
        const uint32_t n = S.counts[HK_HIT_COUNTER(q)];
        const uint32_t n_round = (n + 31u) & ~31u;     // whole warps iterate together: the cooperative light-BVH descent pairs up the lanes of a warp
        const uint32_t* __restrict__ queue = S.q_hit[q];
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
            bool push_lbvh = false;
            uint32_t slot = 0;
            if (i < n) {
            slot = queue[i];
            const float4 hr = S.hit[slot];
            const uint32_t prim0 = HK_HIT_PRIM1(__float_as_uint(hr.y)) - 1u;
            const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
            const float3 o = f3(ra.x, ra.y, ra.z), d = f3(ra.w, rb.x, rb.y);
            const Surf sf = surface_at(D, prim0, hr.z, hr.w, o, d, hr.x);
            const float4 lam = S.lambda[slot];
            const uint32_t fl = S.flags[slot];
            const int depth = HK_FLAG_DEPTH(fl);
            // ---- HandleEmissiveIntersection ----------------------------------------------------------------
            if (sf.arealight > 0u) {
                const float3 wo = -d;
                Spec Le = arealight_Le(D.T, D.lights[sf.arealight - 1], wo, sf.n, lam);
                if (!sp_black(Le)) {
                    const Spec beta = S.beta[slot], r_u = S.r_u[slot], r_l = S.r_l[slot];
                    Spec contrib = beta * Le, fin;
                    if (depth == 0 || (fl & HK_FLAG_SPEC)) fin = contrib / sp_avg(r_u);
                    else {
                        float lcp = bvh_light_pmf(LC, sf.pi, sf.n, (int)sf.arealight);
                        float ct = fabsf(dot3(sf.n, norm3(d)));
                        float lpdf = (ct > 0.0f && sf.area > 0.0f) ? lcp * ((hr.x * hr.x) / (ct * sf.area)) : 0.0f;
                        float den = sp_avg(r_u + r_l * lpdf);
                        fin = den > 1.0e-10f ? contrib / den : contrib / sp_avg(r_u);
                    }
                    S.L[slot] = S.L[slot] + fin;
                }
            }
            // ---- the light sample of next-event estimation (surface-eval.jl:250-342 up to the BSDF evaluation) --------------
            const uint32_t pix = slot % A.n_pixels;
            const int px = (int)(pix % (uint32_t)D.width) + 1, py = (int)(pix / (uint32_t)D.width) + 1;
            const int sidx = slot_sample_idx(A, slot);
            const int bdim = 6 + 7 * depth;
            float4 rec_b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float direct_uc = zsobol_1d(D.sobol, px, py, sidx, bdim + 1, HK_SOBOL_SLOT_BOUNCE(depth, 0), pix);
            float pmf;
#if HK_LIGHTS_COMPACT
            // hits whose sample falls on the light BVH are queued for k_hit_lights_bvh, where every lane of a warp descends (here only
            // the fraction 1 / (n_infinite + 1) of a warp's lanes would: ncu on C3, 13 of 32 lanes in the descent): the hit point, the
            // shading normal and the remapped sample / tree probability wait in the vertex's light-sample record
            bool need; float ub, tree_p;
            const int li = light_select_prologue(LC, direct_uc, pmf, need, ub, tree_p);
            if (need) { S.nee_a[slot] = make_float4(sf.pi.x, sf.pi.y, sf.pi.z, ub); S.nee_c[slot] = make_float4(sf.ns.x, sf.ns.y, sf.ns.z, tree_p); }
            push_lbvh = need;
#else
            const int li = bvh_sample_light_coop(LC, sf.pi, sf.ns, direct_uc, pmf);
#endif
            if (li >= 1 && li <= D.n_lights && pmf > 0.0f) {
                const float2 direct_u = zsobol_2d(D.sobol, px, py, sidx, bdim + 3, HK_SOBOL_SLOT_BOUNCE(depth, 1), pix);
                const LightSample ls = sample_light(LC, D.lights[li - 1], sf.pi, lam, direct_u);
                if (ls.pdf > 0.0f && !sp_black(ls.Li)) {
                    rec_b = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.pdf);
                    S.nee_a[slot] = ls.Li;
                    S.nee_c[slot] = make_float4(ls.p_light.x, ls.p_light.y, ls.p_light.z, ls.delta ? -pmf : pmf);
                }
            }
            S.nee_b[slot] = rec_b;
            }
#if HK_LIGHTS_COMPACT
            warp_push1(S.counts + HK_C_LBVH, S.q_lbvh, push_lbvh, slot);
#else
            (void)push_lbvh;
#endif
        }
    }
}
#if HK_LIGHTS_COMPACT
// The light-BVH descents of a bounce, compacted: every lane of a warp owns one queued hit, so the pair-of-lanes descent runs with all
// 16 pairs busy, twice per 32 hits (bvh_descend_coop).  Then sample_light for the light that was picked, as k_hit_lights does for the
// infinite lights.  Same functions on the same operands: same bits.
__global__ void __launch_bounds__(128, 6) k_hit_lights_bvh(const __grid_constant__ DevScene D, PathState S, PassArgs A) {
    LightCtx LC = light_ctx(D);
    const uint32_t n = S.counts[HK_C_LBVH];
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint32_t slot = 0;
        float3 p = f3(0.0f, 0.0f, 0.0f), ns = f3(0.0f, 0.0f, 0.0f);
        float ub = 0.0f, tree_p = 0.0f;
        if (valid) {
            slot = S.q_lbvh[i];
            const float4 a = S.nee_a[slot], c = S.nee_c[slot];
            p = f3(a.x, a.y, a.z); ub = a.w; ns = f3(c.x, c.y, c.z); tree_p = c.w;
        }
        float pmf = 0.0f;
        const int li = bvh_descend_coop(LC, p, ns, valid, ub, tree_p, 0, pmf);
        if (valid) {
            float4 rec_b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (li >= 1 && li <= D.n_lights && pmf > 0.0f) {
                const uint32_t pix = slot % A.n_pixels;
                const int px = (int)(pix % (uint32_t)D.width) + 1, py = (int)(pix / (uint32_t)D.width) + 1;
                const int sidx = slot_sample_idx(A, slot);
                const int depth = HK_FLAG_DEPTH(S.flags[slot]);
                const int bdim = 6 + 7 * depth;
                const float2 direct_u = zsobol_2d(D.sobol, px, py, sidx, bdim + 3, HK_SOBOL_SLOT_BOUNCE(depth, 1), pix);
                const LightSample ls = sample_light(LC, D.lights[li - 1], p, S.lambda[slot], direct_u);
                if (ls.pdf > 0.0f && !sp_black(ls.Li)) {
                    rec_b = make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.pdf);
                    S.nee_a[slot] = ls.Li;
                    S.nee_c[slot] = make_float4(ls.p_light.x, ls.p_light.y, ls.p_light.z, ls.delta ? -pmf : pmf);
                }
            }
            S.nee_b[slot] = rec_b;
        }
    }
}
#endif
#endif  // HK_TU_LIGHTS
#ifdef HK_TU_SHADE
// Shading of one material type: emissive-hit MIS (surface-eval.jl:147-220), NEE (:250-342 + lights.jl:535-600) and
// BSDF sampling / Russian roulette / continuation ray (:396-512), fused into one kernel per material type.
// resident blocks per SM the shading kernels are compiled for: 4 (<= 128 registers; above that only 3 blocks fit and the
// measured throughput drops 8 %); the coated-diffuse random walk is long enough to prefer 6 blocks even with spills
#ifndef HK_SHADE_MIN_BLOCKS
#define HK_SHADE_MIN_BLOCKS 4
#endif
// SPLIT: emissive-hit MIS and the NEE light sample of the vertex were done by k_hit_lights (scenes whose light BVH is deep enough
// to need the cooperative descent, DevScene::split_lights); otherwise they are done here with the plain serial descent, as the
// light work is then a few hundred instructions and a separate kernel plus its 48-byte record per hit costs more than it saves
// (measured on B200: C3, 10 002 lights, shading 10.4 -> 6.2 ms per 4K sample when split; C2, 3 lights, 1.50 -> 1.67 ms).
// PART: 0 = the whole vertex in one kernel; 1 = emissive-hit MIS + next-event estimation only; 2 = BSDF sampling + continuation only.
// The LayeredBxDF materials run as two kernels (1 then 2, same stream): their eval / pdf walks and their sampling walk are ~100 KB of
// code each, and with both in one kernel the resident warps -- spread over all of it -- miss the instruction cache on almost every
// fetch (ncu on C5: stall_no_instruction 45 warp-cycles per issue, issue slots 11 % busy).  Part 1 only reads what part 2 rewrites.
#ifndef HK_SHADE_TWO_PARTS
#define HK_SHADE_TWO_PARTS 1
#endif
#ifndef HK_SHADE_LAYERED_EXTRA_BLOCKS
#define HK_SHADE_LAYERED_EXTRA_BLOCKS 2
#endif
// 1: the blocks of the coated / layered shading kernels take their items in step (one barrier per item).  Those kernels are 100-260 KB
// of SASS; ncu on C5's coated-diffuse kernel: 18 warps per issue stalled in no_instruction, issue slots 22 % busy -- every warp streams the
// whole kernel through the instruction cache on its own.  With a block's four warps starting each item together they fetch the same lines
// at about the same time: C5 shading 11.89 -> 10.85 ms/step (297 -> 310 Msamples/s).  Work per item and results are unchanged.
#ifndef HK_SHADE_SYNC
#define HK_SHADE_SYNC 1
#endif
#define HK_SHADE_SYNC_TYPE(T) ((T) == HK_MAT_COATED_DIFFUSE || (T) == HK_MAT_COATED_DIFFUSE_TRANSMISSION || (T) == HK_MAT_COATED_CONDUCTOR)
#define HK_SHADE_IS_LAYERED(T) ((T) == HK_MAT_COATED_DIFFUSE || (T) == HK_MAT_COATED_DIFFUSE_TRANSMISSION)
// TEX: some material of this class has textured parameters (HkMaterial.tex / ftex): this instantiation resolves them per hit into a
// local copy of the material; classes without textured materials run the lean instantiation, which has no such code
template <int TYPE, bool SPLIT, int PART, bool TEX>
__global__ void __launch_bounds__(128, HK_SHADE_IS_LAYERED(TYPE) ? HK_SHADE_MIN_BLOCKS + HK_SHADE_LAYERED_EXTRA_BLOCKS : HK_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ DevScene D, PathState S, PassArgs A, int next, int par) {
    const uint32_t n = S.counts[HK_HIT_COUNTER(HK_TYPE_QUEUE(TYPE))];
    MatCtx MC = mat_ctx(D);
    LightCtx LC = light_ctx(D);
    // whole warps iterate together so the aggregated pushes stay converged; HK_SHADE_SYNC: whole blocks do, with a barrier per item
    const bool block_sync = HK_SHADE_SYNC && HK_SHADE_SYNC_TYPE(TYPE);
    const uint32_t n_round = block_sync ? ((n + 127u) & ~127u) : ((n + 31u) & ~31u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        if (block_sync) __syncthreads();
        bool push_shadow = false, push_ray = false;
        uint32_t slot = 0;
        if (i < n) {
            slot = S.q_hit[HK_TYPE_QUEUE(TYPE)][i];
            const float4 hr = S.hit[slot];
            const uint32_t prim0 = HK_HIT_PRIM1(__float_as_uint(hr.y)) - 1u;
            const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
            const float3 o = f3(ra.x, ra.y, ra.z), d = f3(ra.w, rb.x, rb.y);
            const Surf sf = surface_at(D, prim0, hr.z, hr.w, o, d, hr.x);
            const HkMediumInterface mi = D.interfaces[sf.iface - 1];
            const HkMaterial& gmat = D.materials[(D.materials[mi.material - 1].type == HK_MAT_MIX ? S.res_mat[slot] : mi.material) - 1];
            HkMaterial lmat;                                   // TEX: per-hit copy with the textured parameters resolved (only when there are any)
            if (TEX) { MC.local = material_has_textures(gmat); if (MC.local) resolve_material_textures(D, gmat, prim0, hr.z, hr.w, lmat); }
            const HkMaterial& mat = (TEX && MC.local) ? lmat : gmat;
            const float4 lam = S.lambda[slot];
            Spec kd_tex = sp(0.0f);
            if (TYPE == HK_SHADE_MATTE_TEX) kd_tex = textured_kd(D, mat, prim0, hr.z, hr.w, lam);
            const Spec beta = S.beta[slot], r_u = S.r_u[slot], r_l = S.r_l[slot];
            const uint32_t fl = S.flags[slot];
            const int depth = HK_FLAG_DEPTH(fl);
            const uint32_t cur_medium = HK_FLAG_MEDIUM(fl);
            const float3 wo = -d;
            // ---- HandleEmissiveIntersection (SPLIT: done by k_hit_lights) ---------------------------------------
            if (PART != 2 && !SPLIT && sf.arealight > 0u) {
                Spec Le = arealight_Le(D.T, D.lights[sf.arealight - 1], wo, sf.n, lam);
                if (!sp_black(Le)) {
                    Spec contrib = beta * Le, fin;
                    if (depth == 0 || (fl & HK_FLAG_SPEC)) fin = contrib / sp_avg(r_u);
                    else {
                        float lcp = bvh_light_pmf(LC, sf.pi, sf.n, (int)sf.arealight);
                        float ct = fabsf(dot3(sf.n, norm3(d)));
                        float lpdf = (ct > 0.0f && sf.area > 0.0f) ? lcp * ((hr.x * hr.x) / (ct * sf.area)) : 0.0f;
                        float den = sp_avg(r_u + r_l * lpdf);
                        fin = den > 1.0e-10f ? contrib / den : contrib / sp_avg(r_u);
                    }
                    S.L[slot] = S.L[slot] + fin;
                }
            }
            // ---- per-bounce Sobol dimensions (volpath.jl:253-262) ---------------------------------------------
            const uint32_t pix = slot % A.n_pixels;
            const int px = (int)(pix % (uint32_t)D.width) + 1, py = (int)(pix / (uint32_t)D.width) + 1;
            const int sidx = slot_sample_idx(A, slot);
            const int bdim = 6 + 7 * depth;
            // ---- next-event estimation: the light sample comes from k_hit_lights' record (SPLIT) or is drawn here ------------
            if (PART != 2 && D.n_lights > 0) {
                float3 lwi = f3(0.0f, 0.0f, 0.0f), lp = f3(0.0f, 0.0f, 0.0f);
                float lpdf = 0.0f, pmf = 0.0f; bool ldelta = false;
                Spec Li = sp(0.0f);
                if (SPLIT) {
                    const float4 nb4 = S.nee_b[slot];                  // wi, pdf (0 = no usable light sample)
                    if (nb4.w > 0.0f) {
                        const float4 nc4 = S.nee_c[slot];              // p_light, pmf (sign bit set = delta light)
                        lwi = f3(nb4.x, nb4.y, nb4.z); lpdf = nb4.w; lp = f3(nc4.x, nc4.y, nc4.z);
                        ldelta = (__float_as_uint(nc4.w) >> 31) != 0u; pmf = fabsf(nc4.w);
                        Li = S.nee_a[slot];
                    }
                } else {
                    float direct_uc = zsobol_1d(D.sobol, px, py, sidx, bdim + 1, HK_SOBOL_SLOT_BOUNCE(depth, 0), pix);
                    float pm;
                    int li = bvh_sample_light(LC, sf.pi, sf.ns, direct_uc, pm);
                    if (li >= 1 && li <= D.n_lights && pm > 0.0f) {
                        float2 direct_u = zsobol_2d(D.sobol, px, py, sidx, bdim + 3, HK_SOBOL_SLOT_BOUNCE(depth, 1), pix);
                        LightSample ls = sample_light(LC, D.lights[li - 1], sf.pi, lam, direct_u);
                        if (ls.pdf > 0.0f && !sp_black(ls.Li)) { lwi = ls.wi; lpdf = ls.pdf; lp = ls.p_light; ldelta = ls.delta; pmf = pm; Li = ls.Li; }
                    }
                }
                if (lpdf > 0.0f) {
                    BsdfEval be = TYPE == HK_SHADE_MATTE_TEX ? eval_matte_kd(kd_tex, wo, lwi, sf.ns) : eval_bsdf<TYPE>(MC, mat, wo, lwi, sf.ns, lam);
                    if (!sp_black(be.f)) {
                        float ct = fabsf(dot3(lwi, sf.ns));
                        Spec Ld = beta * be.f * Li * ct;
                        if (!sp_black(Ld)) {
                            float3 off = 1.0e-4f * sf.ns;
                            float3 ro = dot3(lwi, sf.ns) > 0.0f ? sf.pi + off : sf.pi - off;
                            float3 tl = lp - ro;
                            float tmax = sqrtf(dot3(tl, tl)) - 1.0e-3f;
                            S.sh_a[slot] = make_float4(ro.x, ro.y, ro.z, lwi.x);
                            S.sh_b[slot] = make_float4(lwi.y, lwi.z, tmax, 0.0f);
                            S.sh_Ld[slot] = Ld;
                            S.sh_ru[slot] = r_u * (ldelta ? 0.0f : be.pdf);
                            S.sh_rl[slot] = r_u * lpdf * pmf;
                            S.sh_medium[slot] = cur_medium;
                            push_shadow = true;
                        }
                    }
                }
            }
            // ---- BSDF sampling + continuation ----------------------------------------------------------------------
            const int new_depth = depth + 1;
            if (PART != 1 && new_depth < D.max_depth) {
                float indirect_uc = zsobol_1d(D.sobol, px, py, sidx, bdim + 4, HK_SOBOL_SLOT_BOUNCE(depth, 2), pix);
                float2 indirect_u = zsobol_2d(D.sobol, px, py, sidx, bdim + 6, HK_SOBOL_SLOT_BOUNCE(depth, 3), pix);
                const bool reg = D.regularize && (fl & HK_FLAG_ANYNS);
                BsdfSample bs = TYPE == HK_SHADE_MATTE_TEX ? sample_matte_kd(kd_tex, mat.f[0], wo, sf.ns, indirect_u)
                                                           : sample_bsdf<TYPE>(MC, mat, wo, sf.ns, lam, indirect_u, indirect_uc, reg);
                if (bs.pdf > 0.0f && !sp_black(bs.f)) {
                    float ct = fabsf(dot3(bs.wi, sf.ns));
                    Spec nb = bs.specular ? beta * bs.f : beta * bs.f * ct / bs.pdf;
                    Spec nrl = bs.specular ? r_u : r_u / bs.pdf;
                    float rr = zsobol_1d(D.sobol, px, py, sidx, bdim + 7, HK_SOBOL_SLOT_BOUNCE(depth, 4), pix);
                    if (russian_roulette(nb, new_depth, rr)) {
                        const float side = dot3(bs.wi, sf.n);
                        uint32_t nm = (mi.inside != mi.outside) ? (side > 0.0f ? mi.outside : mi.inside) : cur_medium;
                        float3 od = side > 0.0f ? sf.n : -sf.n;
                        float3 no = sf.pi + od * 0.0001f;
                        S.ray_a[slot] = make_float4(no.x, no.y, no.z, bs.wi.x);
                        S.ray_b[slot] = make_float4(bs.wi.y, bs.wi.z, HK_INF, 0.0f);
                        S.beta[slot] = nb; S.r_l[slot] = nrl;
                        uint32_t nf = (uint32_t)new_depth | (bs.specular ? HK_FLAG_SPEC : 0u) | (((fl & HK_FLAG_ANYNS) || !bs.specular) ? HK_FLAG_ANYNS : 0u) | (nm << 16);
                        S.flags[slot] = nf;
                        push_ray = true;
                    }
                }
            }
        }
        if (PART == 0) warp_push2(S.counts + HK_CI_SHADOW(par), S.q_shadow, push_shadow, S.counts + HK_C_RAY0 + next, S.q_ray[next], push_ray, slot);
        else if (PART == 1) warp_push1(S.counts + HK_CI_SHADOW(par), S.q_shadow, push_shadow, slot);
        else warp_push1(S.counts + HK_C_RAY0 + next, S.q_ray[next], push_ray, slot);
    }
}
// one shading class = one launch, or two for the LayeredBxDF materials
template <int TYPE, bool TEX> static void launch_shade_class_t(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    if constexpr (HK_SHADE_TWO_PARTS && HK_SHADE_IS_LAYERED(TYPE)) {
        if (D.split_lights) { k_shade<TYPE, true, 1, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par); k_shade<TYPE, true, 2, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par); }
        else { k_shade<TYPE, false, 1, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par); k_shade<TYPE, false, 2, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par); }
    } else {
        if (D.split_lights) k_shade<TYPE, true, 0, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<TYPE, false, 0, TEX><<<grid, 128, 0, st>>>(D, S, A, next, par);
    }
}
template <int TYPE> static void launch_shade_class(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    if ((D.tex_classes >> TYPE) & 1u) launch_shade_class_t<TYPE, true>(grid, st, D, S, A, next, par);
    else launch_shade_class_t<TYPE, false>(grid, st, D, S, A, next, par);
}

#endif  // HK_TU_SHADE
#ifdef HK_TU_MEDIA
// Delta tracking (delta-tracking.jl:142-453) runs as a persistent per-lane-refill loop over the DeltaTracker state machine
// (hk_media.cuh); its per-slot result feeds k_medium_finish = medium NEE + phase-function sampling + routing
// (medium-scatter.jl:15-203), a plain one-thread-per-entry pass where whole warps stay converged.
#ifndef HK_PHASE_MIN
#define HK_PHASE_MIN 16
#endif
#ifndef HK_PHASE_POLICY
#define HK_PHASE_POLICY 0      // 0: per-warp vote between the skip and the event phase; 1: both phases every iteration
#endif
#ifndef HK_MEDIUM_REFILL_MIN
#define HK_MEDIUM_REFILL_MIN HK_PHASE_MIN     // idle lanes needed before the warp fetches new rays (32 = coherent groups, no mid-flight refill)
#endif
#ifndef HK_MEDIA_MIN_BLOCKS
#define HK_MEDIA_MIN_BLOCKS 4      // 128 registers; measured on C4: 1 (150 registers, 3 blocks) -5 %, 5 (102 registers, spills) -11 %
#endif
template <bool RGB>     // RGB: some medium of the scene is an RGBGridMedium (see DeltaTracker::event_step)
__global__ void __launch_bounds__(128, HK_MEDIA_MIN_BLOCKS) k_medium_track(const __grid_constant__ DevScene D, PathState S) {
    extern __shared__ uint32_t s_mask[];
    const uint32_t n = S.counts[HK_C_MEDIUM];
    if (n == 0) return;
    MediaCtx MDC = media_ctx_staged(D, s_mask);
    uint32_t* lc_slot = s_mask + D.smem_mask_words + threadIdx.x;      // this lane's NanoVDB leaf cache (medium_density_cached)
    lc_slot[0] = 0u;
    DeltaTracker T;
    T.in_seg = false;
    bool busy = false, exhausted = false;
#ifdef HK_MEDIA_STATS
    bool had_seg = false;
#endif
    uint32_t slot = 0;
    for (;;) {
        // Phase vote: the warp runs ONE of {refill, event, skip} per iteration, chosen so that the two expensive ones (ray
        // set-up and collision events) only run when at least HK_PHASE_MIN lanes want them or nothing else can progress.
        // (Per-lane refill alone left 3 of 32 lanes active in the event code: at any time only a few lanes sit in a
        // non-empty majorant cell while the rest skip empty ones.)
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        const unsigned ev = __ballot_sync(0xFFFFFFFFu, busy && T.in_seg);
        const unsigned sk = ~(idle | ev);
        if (idle == 0xFFFFFFFFu && exhausted) break;
        if (!exhausted && ((uint32_t)__popc(idle) >= (uint32_t)HK_MEDIUM_REFILL_MIN || (ev | sk) == 0u)) {
            if (!busy) {
                const uint32_t idx = claim_for_idle(S.counts + HK_C_CURSOR_MEDIUM, idle);
                if (idx < n) {
                    slot = S.q_medium[idx];
                    const float4 hr = S.hit[slot];
                    const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
                    const uint32_t fl = S.flags[slot];
                    const float t_max = HK_HIT_PRIM1(__float_as_uint(hr.y)) ? hr.x : HK_INF;
                    T.init(MDC, (int)HK_FLAG_MEDIUM(fl), f3(ra.x, ra.y, ra.z), f3(ra.w, rb.x, rb.y), t_max, S.lambda[slot], S.beta[slot], S.r_u[slot], S.r_l[slot],
                           HK_FLAG_DEPTH(fl), D.max_depth);
                    lc_slot[0] = 0u;      // (another ray, possibly another medium)
                    busy = true;
                }
            }
            exhausted = __any_sync(0xFFFFFFFFu, !busy);
            continue;
        }
        bool fin = false;
#if HK_PHASE_POLICY == 1
        // no vote: lanes outside a segment advance their DDA, then every lane inside a segment (the ones that just entered included)
        // takes one collision event
        if (busy && !T.in_seg) fin = T.skip_step();
        if (busy && !fin && T.in_seg) fin = T.template event_step<RGB>(MDC.T, S.lambda + slot, lc_slot);
#else
#ifdef HK_MEDIA_STATS
        if ((threadIdx.x & 31u) == 0u) { if ((uint32_t)__popc(ev) >= (uint32_t)HK_PHASE_MIN || sk == 0u) { HK_STAT(6, 1); HK_STAT(7, __popc(ev)); } else { HK_STAT(5, 1); HK_STAT(15, __popc(sk)); } }
#endif
        if ((uint32_t)__popc(ev) >= (uint32_t)HK_PHASE_MIN || sk == 0u) {
            if (busy && T.in_seg) {
                fin = T.template event_step<RGB>(MDC.T, S.lambda + slot, lc_slot);
#if HK_PHASE_POLICY == 2
                if (!fin && !T.in_seg) fin = T.skip_step();      // the segment ended: move on to the next one right away (in a cloud it is the neighbouring cell), staying an event lane
#endif
            }
        }
        else if (busy && !T.in_seg) fin = T.skip_step();
#endif
#ifdef HK_MEDIA_STATS
        if (busy && T.in_seg) had_seg = true;
        if (fin) { HK_STAT(had_seg ? 16 : 17, 1); had_seg = false; }
#endif
        if (fin) {
            const DeltaOut& R = T.R;
            if (!sp_black(R.Le_add)) S.L[slot] = S.L[slot] + R.Le_add;
            S.beta[slot] = R.beta; S.r_u[slot] = R.r_u; S.r_l[slot] = R.r_l;
            S.med[slot] = make_float4(R.p.x, R.p.y, R.p.z, R.g);
            S.med_ev[slot] = (uint32_t)R.event;
            busy = false; T.in_seg = false;
        }
    }
}
__global__ void __launch_bounds__(128) k_medium_finish(const __grid_constant__ DevScene D, PathState S, PassArgs A, int next) {
    const uint32_t n = S.counts[HK_C_MEDIUM];
    LightCtx LC = light_ctx(D);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool push_shadow = false, push_ray = false;
        int qid = -1;
        uint32_t slot = 0;
        if (i < n) {
            slot = S.q_medium[i];
            const uint32_t ev = S.med_ev[slot];
            if (ev == HK_EV_SCATTER) {
                const float4 md = S.med[slot];
                const float3 Rp = f3(md.x, md.y, md.z); const float Rg = md.w;
                const Spec Rbeta = S.beta[slot], Rru = S.r_u[slot];
                const float4 ra = S.ray_a[slot], rb = S.ray_b[slot];
                const float3 d = f3(ra.w, rb.x, rb.y);
                const float4 lam = S.lambda[slot];
                const uint32_t fl = S.flags[slot];
                const int depth = HK_FLAG_DEPTH(fl);
                const uint32_t medium = HK_FLAG_MEDIUM(fl);
                const uint32_t pix = slot % A.n_pixels;
                const int px = (int)(pix % (uint32_t)D.width) + 1, py = (int)(pix / (uint32_t)D.width) + 1;
                const int sidx = slot_sample_idx(A, slot);
                const int bdim = 6 + 7 * depth;
                const float3 wo = -d;
                if (D.n_lights > 0) {   // medium_direct_lighting_inner!
                    float pmf;
                    int li = bvh_sample_light_auto(LC, Rp, f3(0, 0, 0), zsobol_1d(D.sobol, px, py, sidx, bdim + 1, HK_SOBOL_SLOT_BOUNCE(depth, 0), pix), pmf);
                    if (li >= 1 && li <= D.n_lights && pmf > 0.0f) {
                        LightSample ls = sample_light(LC, D.lights[li - 1], Rp, lam, zsobol_2d(D.sobol, px, py, sidx, bdim + 3, HK_SOBOL_SLOT_BOUNCE(depth, 1), pix));
                        if (ls.pdf > 0.0f && !sp_black(ls.Li)) {
                            float ph = hg_p(Rg, dot3(wo, ls.wi));
                            if (ph > 0.0f) {
                                float tmax = ls.delta ? len3(ls.p_light - Rp) - 0.001f : 1.0e6f;
                                S.sh_a[slot] = make_float4(Rp.x, Rp.y, Rp.z, ls.wi.x);
                                S.sh_b[slot] = make_float4(ls.wi.y, ls.wi.z, tmax, 0.0f);
                                S.sh_Ld[slot] = Rbeta * ph * ls.Li;
                                S.sh_ru[slot] = Rru * (ls.delta ? 0.0f : ph);
                                S.sh_rl[slot] = Rru * (ls.pdf * pmf);
                                S.sh_medium[slot] = medium;
                                push_shadow = true;
                            }
                        }
                    }
                }
                const int nd = depth + 1;   // medium_scatter_inner!
                if (nd < D.max_depth) {
                    float pdf;
                    float3 wi = sample_hg(Rg, wo, zsobol_2d(D.sobol, px, py, sidx, bdim + 6, HK_SOBOL_SLOT_BOUNCE(depth, 3), pix), pdf);
                    if (pdf > 0.0f) {
                        S.ray_a[slot] = make_float4(Rp.x, Rp.y, Rp.z, wi.x);
                        S.ray_b[slot] = make_float4(wi.y, wi.z, HK_INF, 0.0f);
                        S.r_l[slot] = Rru / pdf;      // beta, r_u already hold the scattered values
                        S.flags[slot] = (uint32_t)nd | HK_FLAG_ANYNS | (medium << 16);
                        push_ray = true;
                    }
                }
            } else if (ev == HK_EV_SURVIVED) {
                const float4 hr = S.hit[slot];
                const uint32_t hb = __float_as_uint(hr.y);
                if (!(sp_black(S.beta[slot]) || sp_black(S.r_u[slot]) || HK_FLAG_DEPTH(S.flags[slot]) >= D.max_depth))
                    qid = HK_HIT_PRIM1(hb) ? hit_queue_id(D, S, slot, hb, hr.x) : HK_C_ESCAPED;
            }
        }
        warp_push(S.counts, nullptr, qid, slot, queue_of(S, qid));
        unsigned hm = __ballot_sync(0xFFFFFFFFu, qid >= HK_C_HIT0);
        if ((threadIdx.x & 31u) == 0 && hm) atomicAdd(S.counts + HK_C_TOTAL_HITS, (uint32_t)__popc(hm));
        { const unsigned vm = hm | __ballot_sync(0xFFFFFFFFu, push_shadow || push_ray);      // surface vertices + medium scatter vertices
          if ((threadIdx.x & 31u) == 0 && vm) atomicAdd(S.path_vertices, (unsigned long long)__popc(vm)); }
        warp_push2(S.counts + HK_C_SHADOW, S.q_shadow, push_shadow, S.counts + HK_C_RAY0 + next, S.q_ray[next], push_ray, slot);
    }
}

#endif  // HK_TU_MEDIA
#ifdef HK_TU_TRACE
// trace_shadow_transmittance + vp_trace_shadow_rays_kernel!, intersection.jl:302-406, 565-600.
// Opaque-only scenes (no interface with inside != outside, no media): visibility is a single any-hit query, which yields
// the same T in {0,1} as the reference's closest-hit loop.  Otherwise the ordered closest-hit walk with ratio tracking.
struct ShadowRayIO {
    const PathState& S;
    HK_DEV uint32_t load(uint32_t idx, float3& o, float3& d, float& tm) const {
        const uint32_t slot = S.q_shadow[idx];
        const float4 sa = S.sh_a[slot], sb = S.sh_b[slot];
        o = f3(sa.x, sa.y, sa.z); d = f3(sa.w, sb.x, sb.y);
        tm = sb.z < 1.0e-6f ? -1.0f : sb.z;          // t_remaining < 1e-6: the reference's loop breaks => treated as occluded
        return slot;
    }
    HK_DEV void store(uint32_t slot, const HitRec& h) const {
        if (h.prim1 != 0u || S.sh_b[slot].z < 1.0e-6f) return;      // blocked
        const float den = sp_avg(S.sh_ru[slot] + S.sh_rl[slot]);     // T = tr_u = tr_l = 1
        if (den > 1.0e-10f) {
            const Spec fin = S.sh_Ld[slot] / den;
            if (!sp_black(fin)) S.L[slot] = S.L[slot] + fin;
        }
    }
};
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(HK_TRACE_THREADS, INST ? HK_TRACE_BLOCKS_PER_SM_INST : HK_TRACE_BLOCKS_PER_SM) k_shadow_opaque(const __grid_constant__ DevScene D, PathState S, unsigned long long* work, int par) {
    HK_TRACE_SMEM(INST);
    // reference quirk (volpath.jl:571-609): shadow rays are only traced inside the `n_hits > 0` branch
    if (S.counts[HK_CI_TOTAL_HITS(par)] == 0) return;
    const uint32_t n = S.counts[HK_CI_SHADOW(par)];
    uint32_t traced = 0, wn = 0, wt = 0;
    ShadowRayIO io{S};
    trace_queue<true, COUNT, INST>(D.bvh, sm_stack + threadIdx.x, sm_wray + threadIdx.x, n, S.counts + HK_CI_CURSOR_SHADOW(par), io, traced, wn, wt);
    count_rays(S.rays_traced, traced);
    if (COUNT) { count_rays(work + 3, traced); count_rays(work + 4, wn); count_rays(work + 5, wt); }
}
// Scenes with media / medium interfaces: trace_shadow_transmittance (intersection.jl:302-406) walks the shadow ray segment
// by segment -- closest hit, ratio tracking through the current medium up to it, cross the boundary, repeat (<= 10 times).
// Here every segment is one wavefront ROUND of two persistent kernels over the still-unresolved shadow rays:
//   k_shadow_seg_trace : closest hit of the segment, through the same per-lane-refill traversal loop as k_trace;
//   k_shadow_seg_ratio : ratio tracking as a per-lane-refill loop over the RatioTracker state machine, then resolve
//                        (visible -> add the contribution, blocked -> drop, boundary -> next round's queue).
// The one-thread-per-ray-to-completion form this replaces ran with 4.5 of 32 lanes active on C4 (ncu, profiles/).
struct ShadowSegIO {
    const PathState& S; const uint32_t* __restrict__ q;
    HK_DEV uint32_t load(uint32_t idx, float3& o, float3& d, float& tm) const {
        const uint32_t slot = q[idx];
        const float4 sa = S.sh_a[slot], sb = S.sh_b[slot];
        o = f3(sa.x, sa.y, sa.z); d = f3(sa.w, sb.x, sb.y);
        tm = sb.z < 1.0e-6f ? -1.0f : sb.z;          // t_remaining < 1e-6: the reference's loop breaks (ray dropped by the ratio pass)
        return slot;
    }
    HK_DEV void store(uint32_t slot, const HitRec& h) const { S.sh_hit[slot] = make_float4(h.t, __uint_as_float(h.prim1), h.b1, h.b2); }
};
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(HK_TRACE_THREADS, INST ? HK_TRACE_BLOCKS_PER_SM_INST : HK_TRACE_BLOCKS_PER_SM) k_shadow_seg_trace(const __grid_constant__ DevScene D, PathState S, int round, unsigned long long* work) {
    HK_TRACE_SMEM(INST);
    if (S.counts[HK_C_TOTAL_HITS] == 0) return;      // reference quirk (volpath.jl:571-609): shadow pass only inside `n_hits > 0`
    const uint32_t n = S.counts[round == 0 ? HK_C_SHADOW : HK_C_SHROUND0 + round];
    uint32_t traced = 0, wn = 0, wt = 0;
    ShadowSegIO io{S, (round & 1) ? S.q_shadow2 : S.q_shadow};
    trace_queue<false, COUNT, INST>(D.bvh, sm_stack + threadIdx.x, sm_wray + threadIdx.x, n, S.counts + HK_C_SHCUR_TRACE + round, io, traced, wn, wt);
    count_rays(S.rays_traced, traced);
    if (COUNT) { count_rays(work + 3, traced); count_rays(work + 4, wn); count_rays(work + 5, wt); }
}
#endif  // HK_TU_TRACE
#ifdef HK_TU_MEDIA
template <bool RGB>
__global__ void __launch_bounds__(128, HK_MEDIA_MIN_BLOCKS) k_shadow_seg_ratio(const __grid_constant__ DevScene D, PathState S, int round) {
    if (S.counts[HK_C_TOTAL_HITS] == 0) return;
    const uint32_t n = S.counts[round == 0 ? HK_C_SHADOW : HK_C_SHROUND0 + round];
    const uint32_t* __restrict__ q = (round & 1) ? S.q_shadow2 : S.q_shadow;
    uint32_t* q_next = (round & 1) ? S.q_shadow : S.q_shadow2;
    if (n == 0) return;
    extern __shared__ uint32_t s_mask[];
    MediaCtx MDC = media_ctx_staged(D, s_mask);
    uint32_t* lc_slot = s_mask + D.smem_mask_words + threadIdx.x;
    lc_slot[0] = 0u;
    RatioTracker R;
    R.in_seg = false;
    bool busy = false, exhausted = false, tracking = false, apass = false;
#ifdef HK_MEDIA_STATS
    bool had_seg = false;
#endif
    uint32_t slot = 0;
    for (;;) {
        // same phase vote as k_medium_track; a lane whose segment needs no tracking (vacuum) resolves in the event phase
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        const unsigned ev = __ballot_sync(0xFFFFFFFFu, busy && (!tracking || R.in_seg));
        const unsigned sk = ~(idle | ev);
        if (idle == 0xFFFFFFFFu && exhausted) break;
        if (!exhausted && ((uint32_t)__popc(idle) >= (uint32_t)HK_MEDIUM_REFILL_MIN || (ev | sk) == 0u)) {
            bool past_end = false;      // only a claim beyond the queue's end means the queue is exhausted: a claimed ray that is dropped
                                        // below (blocked by an opaque surface, t_remaining < 1e-6) leaves its lane idle with work still queued
            if (!busy) {
                const uint32_t idx = claim_for_idle(S.counts + HK_C_SHCUR_RATIO + round, idle);
                past_end = idx >= n;
                if (idx < n) {
                    slot = q[idx];
                    const float4 sa = S.sh_a[slot], sb = S.sh_b[slot];
                    const float4 h = S.sh_hit[slot];
                    const uint32_t hp = HK_HIT_PRIM1(__float_as_uint(h.y));
                    bool opaque = false;                          // an opaque surface (alpha == 1) blocks: no tracking needed
                    apass = false;                                // ... unless the stochastic alpha test lets the ray through (intersection.jl:349-372)
                    if (hp != 0u) {
                        const HkMediumInterface mi = D.interfaces[prim_iface(D, hp - 1u) - 1]; opaque = mi.inside == mi.outside;
                        if (opaque && D.has_alpha && !(sb.z < 1.0e-6f) && HK_HIT_MTYPE(__float_as_uint(h.y)) == HK_SHADE_MATTE_TEX) {
                            apass = alpha_skips(D, hp - 1u, h.z, h.w, f3(sa.x, sa.y, sa.z), f3(sa.w, sb.x, sb.y)); opaque = !apass;
                        }
                    }
                    if (!(sb.z < 1.0e-6f) && !opaque) {           // t_rem < 1e-6: dropped (the reference's loop ends without a visible ray)
                        busy = true;
                        const uint32_t cur = S.sh_medium[slot];
                        tracking = cur != 0u;
                        R.in_seg = false;
                        if (tracking) {
                            const float t_seg = hp ? h.x : sb.z;
                            R.init(MDC, (int)cur, f3(sa.x, sa.y, sa.z), f3(sa.w, sb.x, sb.y), t_seg, S.lambda[slot]);
                            lc_slot[0] = 0u;
                        }
                    }
                }
            }
            exhausted = __any_sync(0xFFFFFFFFu, past_end);
            continue;
        }
        bool fin = false;
#if HK_PHASE_POLICY == 1
        if (busy && tracking && !R.in_seg) fin = R.skip_step();
        if (busy && !fin) { if (!tracking) fin = true; else if (R.in_seg) fin = R.template event_step<RGB>(MDC.T, S.lambda + slot, lc_slot); }
#else
#ifdef HK_MEDIA_STATS
        if ((threadIdx.x & 31u) == 0u) { if ((uint32_t)__popc(ev) >= (uint32_t)HK_PHASE_MIN || sk == 0u) { HK_STAT(14, 1); HK_STAT(13, __popc(ev)); } }
#endif
        if ((uint32_t)__popc(ev) >= (uint32_t)HK_PHASE_MIN || sk == 0u) {
            if (busy) {
                if (!tracking) fin = true;
                else if (R.in_seg) {
                    fin = R.template event_step<RGB>(MDC.T, S.lambda + slot, lc_slot);
#if HK_PHASE_POLICY == 2
                    if (!fin && !R.in_seg) fin = R.skip_step();
#endif
                }
            }
        }
        else if (busy && tracking && !R.in_seg) fin = R.skip_step();
#endif
#ifdef HK_MEDIA_STATS
        if (busy && tracking && R.in_seg) had_seg = true;
        if (fin) { HK_STAT(!tracking ? 20 : (had_seg ? 18 : 19), 1); had_seg = false; }
#endif
        if (fin) {
            // ---- the segment is done: fold its transmittance in and resolve ------------------------------------------
            Spec T = sp(1.0f), tu = sp(1.0f), tl = sp(1.0f);
            if (round > 0) { T = S.sh_T[slot]; tu = S.sh_tu[slot]; tl = S.sh_tl[slot]; }
            if (tracking) { T = T * R.T_ray; tu = tu * R.r_u; tl = tl * R.r_l; }
            const float4 h = S.sh_hit[slot];
            const uint32_t prim1 = HK_HIT_PRIM1(__float_as_uint(h.y));
            bool visible = false;
            if (prim1 == 0u) visible = true;
            else {
                const uint32_t prim0 = prim1 - 1u;
                const HkMediumInterface mi = D.interfaces[prim_iface(D, prim0) - 1];
                if (mi.inside != mi.outside || apass) {            // (an opaque surface, alpha == 1, blocks: nothing to do)
                    if (!apass && sp_black(T)) visible = true;     // reference: leaves the loop "visible" with T == 0 -> contributes nothing
                    else if (round + 1 < HK_SHADOW_ROUNDS) {       // (an alpha pass-through continues whatever T is, medium unchanged)
                        const float4 sa = S.sh_a[slot], sb = S.sh_b[slot];
                        const float3 o = f3(sa.x, sa.y, sa.z), d = f3(sa.w, sb.x, sb.y);
                        const bool entering = dot3(d, geometric_normal(D, prim0)) < 0.0f;
                        const float3 no = o + d * (h.x + 1.0e-4f);
                        if (!apass) S.sh_medium[slot] = entering ? mi.inside : mi.outside;
                        S.sh_a[slot] = make_float4(no.x, no.y, no.z, d.x);
                        S.sh_b[slot] = make_float4(sb.x, sb.y, sb.z - h.x - 1.0e-4f, 0.0f);
                        S.sh_T[slot] = T; S.sh_tu[slot] = tu; S.sh_tl[slot] = tl;
                        push_active(S.counts + HK_C_SHROUND0 + round + 1, q_next, slot);
                    }
                }
            }
            if (visible && !sp_black(T)) {
                const float den = sp_avg(S.sh_ru[slot] * tu + S.sh_rl[slot] * tl);
                if (den > 1.0e-10f) {
                    const Spec contrib = S.sh_Ld[slot] * T / den;
                    if (!sp_black(contrib)) S.L[slot] = S.L[slot] + contrib;
                }
            }
            busy = false;
        }
    }
}

#endif  // HK_TU_MEDIA
#ifdef HK_TU_CORE
// vp_accumulate_to_rgb_kernel!, volpath.jl:326-375.  One thread per pixel walks the batch in sample order, so the
// f32 sums see the samples in exactly the order the reference's per-sample passes do (and no atomics are needed).
__global__ void __launch_bounds__(256) k_film_accumulate(const __grid_constant__ DevScene D, PathState S, PassArgs A) {
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < A.n_pixels; pix += gridDim.x * blockDim.x) {
        float r = S.pixel_rgb[3 * (size_t)pix], g = S.pixel_rgb[3 * (size_t)pix + 1], b = S.pixel_rgb[3 * (size_t)pix + 2], ws = S.pixel_weight[pix];
        for (int k = 0; k < A.n_batch; k++) {
            const size_t slot = (size_t)k * A.n_pixels + pix;
            float3 rgb = xyz_to_linear_srgb(spectral_to_xyz(D.T, S.L[slot], S.lambda[slot], S.lpdf[slot]));
            rgb = f3(rgb.x != rgb.x ? rgb.x : fmaxf(0.0f, rgb.x), rgb.y != rgb.y ? rgb.y : fmaxf(0.0f, rgb.y), rgb.z != rgb.z ? rgb.z : fmaxf(0.0f, rgb.z));
            float m = fmaxf(fmaxf(rgb.x, rgb.y), rgb.z);
            if (m > D.max_component_value) rgb = rgb * (D.max_component_value / m);
            float w = S.fweight[slot];
            r += w * rgb.x; g += w * rgb.y; b += w * rgb.z; ws += w;
        }
        S.pixel_rgb[3 * (size_t)pix] = r; S.pixel_rgb[3 * (size_t)pix + 1] = g; S.pixel_rgb[3 * (size_t)pix + 2] = b; S.pixel_weight[pix] = ws;
    }
}

// vp_finalize_film_kernel!, volpath.jl:384-417: framebuffer[py, px] in (H, W) column-major
__global__ void __launch_bounds__(256) k_film_finalize(const float* __restrict__ rgb, const float* __restrict__ wsum, float* __restrict__ out, int W, int H) {
    const uint32_t n = (uint32_t)W * (uint32_t)H;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += gridDim.x * blockDim.x) {
        const uint32_t px = pix % (uint32_t)W, py = pix / (uint32_t)W;
        const float w = wsum[pix];
        float r = 0.0f, g = 0.0f, b = 0.0f;
        if (w > 0.0f) { float inv = 1.0f / w; r = rgb[3 * (size_t)pix] * inv; g = rgb[3 * (size_t)pix + 1] * inv; b = rgb[3 * (size_t)pix + 2] * inv; }
        float* o = out + ((size_t)px * H + py) * 3;
        o[0] = r; o[1] = g; o[2] = b;
    }
}


// ---- postprocess!, src/postprocess.jl:55-182, 187-230 (per pixel) --------------------------------------------------------
HK_DEV float pp_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
HK_DEV float pp_uncharted2(float x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
HK_DEV float pp_filmic(float x) { x = fmaxf(0.0f, x - 0.004f); return (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f); }
HK_DEV float pp_aces(float x) { return pp_clamp01((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f)); }
HK_DEV float3 postprocess_pixel(const HkPostprocess& P, float3 c) {
    float r = c.x * P.exposure, g = c.y * P.exposure, b = c.z * P.exposure;
    if (P.apply_wb) {
        const float ro = P.wb[0] * r + P.wb[1] * g + P.wb[2] * b, go = P.wb[3] * r + P.wb[4] * g + P.wb[5] * b, bo = P.wb[6] * r + P.wb[7] * g + P.wb[8] * b;
        r = fmaxf(0.0f, ro); g = fmaxf(0.0f, go); b = fmaxf(0.0f, bo);
    }
    r = r * P.imaging_ratio; g = g * P.imaging_ratio; b = b * P.imaging_ratio;
    switch (P.tonemap_mode) {
        case HK_TONEMAP_REINHARD: {
            const float lum = 0.2126f * r + 0.7152f * g + 0.0722f * b;
            const float s = lum > 0.0f ? 1.0f / (1.0f + lum) : 1.0f;
            r = pp_clamp01(r * s); g = pp_clamp01(g * s); b = pp_clamp01(b * s); break; }
        case HK_TONEMAP_REINHARD_EXT: {
            const float lum = 0.2126f * r + 0.7152f * g + 0.0722f * b;
            const float lw2 = P.white_point * P.white_point;
            const float s = lum > 0.0f ? (1.0f + lum / lw2) / (1.0f + lum) : 1.0f;
            r = pp_clamp01(r * s); g = pp_clamp01(g * s); b = pp_clamp01(b * s); break; }
        case HK_TONEMAP_ACES: r = pp_aces(r); g = pp_aces(g); b = pp_aces(b); break;
        case HK_TONEMAP_UNCHARTED2: {
            const float ws = 1.0f / pp_uncharted2(11.2f);
            r = pp_clamp01(pp_uncharted2(r * 2.0f) * ws); g = pp_clamp01(pp_uncharted2(g * 2.0f) * ws); b = pp_clamp01(pp_uncharted2(b * 2.0f) * ws); break; }
        case HK_TONEMAP_FILMIC: r = pp_filmic(r); g = pp_filmic(g); b = pp_filmic(b); break;
        default: r = pp_clamp01(r); g = pp_clamp01(g); b = pp_clamp01(b); break;
    }
    if (P.apply_gamma) { r = dm_powf(r, P.inv_gamma); g = dm_powf(g, P.inv_gamma); b = dm_powf(b, P.inv_gamma); }
    return f3(r, g, b);
}
// framebuffer = sum / weight -> postprocess_pixel, same (H, W) column-major layout as k_film_finalize.
// depth: film.depth for the escaped-ray mask (postprocess.jl:220-245), read with the reference's row flip.
__global__ void __launch_bounds__(256) k_film_postprocess(const float* __restrict__ rgb, const float* __restrict__ wsum, float* __restrict__ out, int W, int H, HkPostprocess P,
                                                           const float* __restrict__ depth) {
    const uint32_t n = (uint32_t)W * (uint32_t)H;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += gridDim.x * blockDim.x) {
        const uint32_t px = pix % (uint32_t)W, py = pix / (uint32_t)W;
        const float w = wsum[pix];
        float3 c = f3(0.0f, 0.0f, 0.0f);
        if (w > 0.0f) { float inv = 1.0f / w; c = f3(rgb[3 * (size_t)pix] * inv, rgb[3 * (size_t)pix + 1] * inv, rgb[3 * (size_t)pix + 2] * inv); }
        c = postprocess_pixel(P, c);
        if (P.mask_escaped) {
            const int d_row = H - (int)py, col = (int)px + 1;      // 1-based (row, col) = (py + 1, px + 1); d_row = H - row + 1
            int escaped = 0, total = 0;
            for (int dr = -1; dr <= 1; dr++)
                for (int dc = -1; dc <= 1; dc++) {
                    const int nr = d_row + dr, nc = col + dc;
                    if (nr >= 1 && nr <= H && nc >= 1 && nc <= W) { escaped += isinf(depth[(size_t)(nc - 1) * H + (nr - 1)]) ? 1 : 0; total++; }
                }
            const float alpha = (float)escaped / (float)total;
            c = f3(c.x * (1.0f - alpha) + P.background[0] * alpha, c.y * (1.0f - alpha) + P.background[1] * alpha, c.z * (1.0f - alpha) + P.background[2] * alpha);
        }
        float* o = out + ((size_t)px * H + py) * 3;
        o[0] = c.x; o[1] = c.y; o[2] = c.z;
    }
}

#endif  // HK_TU_CORE
#ifdef HK_TU_TRACE
// aux_buffer_kernel!, film.jl:433-488: one centre-of-pixel primary ray per pixel; idx runs over the (H, W) column-major buffers
template <bool INST>
__global__ void __launch_bounds__(HK_TRACE_THREADS) k_aux_buffers(const __grid_constant__ DevScene D, float* __restrict__ albedo, float* __restrict__ normal,
                                                                   float* __restrict__ depth, float miss_depth) {
    HK_TRACE_SMEM(INST);
    const uint32_t n = (uint32_t)D.width * (uint32_t)D.height, H = (uint32_t)D.height;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const uint32_t row = idx % H + 1u, col = idx / H + 1u;
        float3 o, d;
        camera_generate_ray(D.camera, (float)col + 0.5f, (float)row + 0.5f, make_float2(0.5f, 0.5f), o, d);
        const HitRec h = bvh8_trace<false, false, INST>(D.bvh, sm_stack + threadIdx.x, sm_wray + threadIdx.x, o, d, HK_INF);
        float3 nn = f3(0.0f, 0.0f, 0.0f); float dep = miss_depth, alb = 0.0f;
        if (h.prim1) {
            const Surf sf = surface_at(D, HK_HIT_PRIM1(h.prim1) - 1u, h.b1, h.b2, o, d, h.t);
            const float3 v = sf.pi - o;
            nn = sf.n; dep = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z); alb = 0.8f;
        }
        albedo[3 * (size_t)idx] = alb; albedo[3 * (size_t)idx + 1] = alb; albedo[3 * (size_t)idx + 2] = alb;
        normal[3 * (size_t)idx] = nn.x; normal[3 * (size_t)idx + 1] = nn.y; normal[3 * (size_t)idx + 2] = nn.z;
        depth[idx] = dep;
    }
}

// stand-alone traversal: rays [n][8] -> hits [n][4]   (hk_trace_closest / hk_trace_any)
template <bool ANY>
struct BatchRayIO {
    const float4* __restrict__ rays; float4* __restrict__ hits; uint8_t* __restrict__ occluded;
    HK_DEV uint32_t load(uint32_t i, float3& o, float3& d, float& tm) const {
        const float4 ra = __ldg(rays + 2 * (size_t)i), rb = __ldg(rays + 2 * (size_t)i + 1);
        o = f3(ra.x, ra.y, ra.z); d = f3(ra.w, rb.x, rb.y); tm = rb.z;
        return i;
    }
    HK_DEV void store(uint32_t i, const HitRec& h) const {
        if (ANY) occluded[i] = h.prim1 ? 1 : 0;
        else hits[i] = make_float4(h.prim1 ? h.t : __ldg(rays + 2 * (size_t)i + 1).z, __uint_as_float(HK_HIT_PRIM1(h.prim1)), h.b1, h.b2);
    }
};
template <bool ANY, bool COUNT, bool INST>
__global__ void __launch_bounds__(HK_TRACE_THREADS, INST ? HK_TRACE_BLOCKS_PER_SM_INST : HK_TRACE_BLOCKS_PER_SM) k_trace_batch(DevBvh B, const float4* __restrict__ rays, uint32_t n, float4* __restrict__ hits,
                                                                uint8_t* __restrict__ occluded, uint32_t* cursor, unsigned long long* counters) {
    HK_TRACE_SMEM(INST);
    uint32_t traced = 0, nn = 0, nt = 0;
    BatchRayIO<ANY> io{rays, hits, occluded};
    trace_queue<ANY, COUNT, INST>(B, sm_stack + threadIdx.x, sm_wray + threadIdx.x, n, cursor, io, traced, nn, nt);
    if (COUNT) { count_rays(counters, nn); count_rays(counters + 1, nt); }
}

// detect_camera_medium, intersection.jl:690-747 (single thread; run once per camera change, not once per sample)
template <bool INST>
__global__ void k_detect_camera_medium(const __grid_constant__ DevScene D, uint32_t* out) {
    HK_TRACE_SMEM(INST);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float3 o = xf_point(D.camera.camera_to_world, f3(0, 0, 0));
    const float3 d = f3(0.57735027f, 0.57735027f, 0.57735027f);
    uint32_t res = 0;
    for (int it = 0; it < 16; it++) {
        HitRec h = bvh8_trace<false, false, INST>(D.bvh, sm_stack, sm_wray, o, d, HK_INF);
        if (h.prim1 == 0) break;
        const uint32_t prim0 = HK_HIT_PRIM1(h.prim1) - 1u;
        const HkMediumInterface mi = D.interfaces[prim_iface(D, prim0) - 1];
        float3 n = geometric_normal(D, prim0);
        if (mi.inside != mi.outside) { res = dot3(-d, n) > 0.0f ? mi.outside : mi.inside; break; }
        float3 pi = o + d * h.t;
        o = pi + (dot3(d, n) > 0.0f ? n : -n) * 1.0e-4f;
    }
    *out = res;
}
#endif  // HK_TU_TRACE
