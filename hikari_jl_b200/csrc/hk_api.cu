// hk_api.cu — C ABI of libhikari_cuda.so (include/hikari_cuda.h): context, uploads, render loop, film, traversal.
// There is NO CPU fallback anywhere in this file: every compute entry point needs a CUDA device and fails with
// HK_ERR_NO_DEVICE / HK_ERR_CUDA otherwise.
#define HK_TU_CORE
#include <cstdlib>
#include "hk_context.h"
#include "hk_launch.h"
#include "hk_denoise.cuh"
#include "hk_bvh.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <array>
#include <vector>
#include <limits>

static int grid_for(const HkContext* c, size_t n, int block, int per_sm) {
    size_t need = (n + block - 1) / block;
    size_t cap = (size_t)c->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

struct StageScope {
    HkContext* c; int stage; HkContext::StageEv* ev = nullptr;
    StageScope(HkContext* ctx, int st) : c(ctx), stage(st) {
        if (!(c->profiling & 1)) return;
        if (c->stage_ev_used == c->stage_events.size()) {
            HkContext::StageEv e; e.stage = st; cudaEventCreate(&e.a); cudaEventCreate(&e.b); c->stage_events.push_back(e);
        }
        ev = &c->stage_events[c->stage_ev_used++]; ev->stage = st;
        cudaEventRecord(ev->a, c->stream);
    }
    ~StageScope() { if (ev) cudaEventRecord(ev->b, c->stream); c->launches++; }
};
static void collect_stage_times(HkContext* c, double* also = nullptr) {
    cudaStreamSynchronize(c->stream);
    for (size_t i = 0; i < c->stage_ev_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->stage_events[i].a, c->stage_events[i].b) == cudaSuccess) {
            c->stage_ms[c->stage_events[i].stage] += ms; c->stage_launches[c->stage_events[i].stage]++;
            if (also) also[c->stage_events[i].stage] += ms;
        }
    }
    c->stage_ev_used = 0;
}

extern "C" {

int32_t hk_abi_version(void) { return HK_ABI_VERSION; }

// The frame pipeline (hk_context.h: HK_N_LANES) keeps ~20 streams busy.  The driver maps streams onto
// CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8); streams that share a queue serialise, and with 8 the lanes
// barely overlap (C3 4K, one sample per call: 808 Msamples/s with 8 queues, 1000 with 32; profiles/r02_e2e_lanes.txt).  The
// variable is read when the CUDA context is created, so it is set when the library is LOADED, unless the host chose a value;
// a host that initialises CUDA before loading the library sets it itself.
__attribute__((constructor)) static void hk_on_load() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

int32_t hk_create(int32_t device, HkContext** out) {
    if (!out) return HK_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return HK_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return HK_ERR_NO_DEVICE;
    HkContext* ctx = new HkContext();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return HK_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    for (auto& s : ctx->shade_streams) if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { s = nullptr; ctx->concurrent_shade = false; }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    for (auto& e : ctx->ev_join) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (std::getenv("HK_SERIAL_SHADE")) ctx->concurrent_shade = false;
    ctx->sort_rays = std::getenv("HK_SORT_RAYS") != nullptr;      // off: measured no gain (profiles/r02_sort_rays_ab.txt)
    if (cudaStreamCreateWithFlags(&ctx->shadow_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->shadow_stream = nullptr;
    cudaEventCreateWithFlags(&ctx->ev_shaded, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_shadowed, cudaEventDisableTiming);
    if (std::getenv("HK_SERIAL_SHADOW") && ctx->shadow_stream) { cudaStreamDestroy(ctx->shadow_stream); ctx->shadow_stream = nullptr; }
    if (ctx->b_counts.alloc(sizeof(uint32_t) * HK_N_COUNTERS + 64) != cudaSuccess || ctx->b_trace_ctr.alloc(64) != cudaSuccess || ctx->b_work_ctr.alloc(64) != cudaSuccess) { delete ctx; return HK_ERR_OOM; }
    cudaMemset(ctx->b_counts.p, 0, ctx->b_counts.bytes); cudaMemset(ctx->b_trace_ctr.p, 0, 64); cudaMemset(ctx->b_work_ctr.p, 0, 64);
    if (ctx->b_scratch_u32.alloc(64) != cudaSuccess) { delete ctx; return HK_ERR_OOM; }
    cudaMemset(ctx->b_scratch_u32.p, 0, 64);
    for (int i = 0; i < HK_N_STAGES; i++) { ctx->stage_ms[i] = 0; ctx->stage_launches[i] = 0; }
    {   // the extra render lanes (frame pipelining); any failure just leaves them off
        bool ok = true;
        for (auto& A : ctx->alts) {
            ok = ok && cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking) == cudaSuccess;
            for (auto& s : A.shade_streams) ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaStreamCreateWithFlags(&A.shadow_stream, cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreate(&A.ev0) == cudaSuccess && cudaEventCreate(&A.ev1) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&A.ev_fork, cudaEventDisableTiming) == cudaSuccess;
            for (auto& e : A.ev_join) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&A.ev_shaded, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&A.ev_shadowed, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && A.b_counts.alloc(sizeof(uint32_t) * HK_N_COUNTERS + 64) == cudaSuccess;
            if (ok) cudaMemset(A.b_counts.p, 0, A.b_counts.bytes);
        }
        for (auto& e : ctx->ev_lane_film) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_film_touch, cudaEventDisableTiming) == cudaSuccess;
        ctx->lanes_ready = ok && ctx->concurrent_shade && ctx->shadow_stream != nullptr && !std::getenv("HK_NO_FRAME_PIPELINE");
        if (!ok) cudaGetLastError();
        if (const char* e = std::getenv("HK_LANE_MAX_COUNT")) ctx->lane_max_count = std::max(1, atoi(e));
    }
    *out = ctx;
    return HK_OK;
}
int32_t hk_destroy(HkContext* ctx) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&ctx->b_sobol, &ctx->b_cie_x, &ctx->b_cie_y, &ctx->b_cie_z, &ctx->b_d65, &ctx->b_rgb_scale, &ctx->b_rgb_coeffs, &ctx->b_nodes, &ctx->b_tris,
                      &ctx->b_esc, &ctx->b_mat_pre, &ctx->b_light_pre, &ctx->b_med_pre, &ctx->b_sobol_top, &ctx->b_sobol_dims, &ctx->b_sobol_dimhash, &ctx->b_pos, &ctx->b_nrm, &ctx->b_idx, &ctx->b_meta, &ctx->b_mats, &ctx->b_ifaces, &ctx->b_spec_l, &ctx->b_spec_v, &ctx->b_spec_o, &ctx->b_lights,
                      &ctx->b_env, &ctx->b_lnodes, &ctx->b_trails, &ctx->b_inf, &ctx->b_media, &ctx->b_f_func, &ctx->b_f_mcdf, &ctx->b_f_mfunc, &ctx->b_f_ccdf,
                      &ctx->b_state, &ctx->b_counts, &ctx->b_rays, &ctx->b_film, &ctx->b_scratch_u32, &ctx->b_trace_ctr, &ctx->b_work_ctr, &ctx->b_readback};
    for (auto& e : ctx->stage_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (DevBuf* b : bufs) b->release();
    for (auto& b : ctx->env_bufs) b.release();
    for (auto& b : ctx->media_bufs) b.release();
    for (auto& b : ctx->mask_bufs) b.release();
    for (auto& b : ctx->dense_bufs) b.release();
    for (auto& b : ctx->tex_bufs) b.release();
    ctx->b_aux.release(); ctx->b_denoise.release(); ctx->b_uvs.release(); ctx->b_textures.release();
    ctx->b_inst_recs.release(); ctx->b_instances.release(); ctx->b_inst_base.release();
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (auto& s : ctx->shade_streams) if (s) cudaStreamDestroy(s);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->shadow_stream) cudaStreamDestroy(ctx->shadow_stream);
    if (ctx->ev_shaded) cudaEventDestroy(ctx->ev_shaded);
    if (ctx->ev_shadowed) cudaEventDestroy(ctx->ev_shadowed);
    for (auto& e : ctx->ev_join) if (e) cudaEventDestroy(e);
    for (int i = 0; i < HK_N_READOUTS; i++) { if (ctx->ev_final[i]) cudaEventDestroy(ctx->ev_final[i]); if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]); ctx->b_readback_async[i].release(); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& A : ctx->alts) {
        A.b_state.release(); A.b_counts.release();
        if (A.stream) cudaStreamDestroy(A.stream);
        for (auto& s2 : A.shade_streams) if (s2) cudaStreamDestroy(s2);
        if (A.shadow_stream) cudaStreamDestroy(A.shadow_stream);
        cudaEvent_t evs[] = {A.ev_fork, A.ev_shaded, A.ev_shadowed, A.ev0, A.ev1};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        for (auto& e : A.ev_join) if (e) cudaEventDestroy(e);
    }
    for (auto& e : ctx->ev_lane_film) if (e) cudaEventDestroy(e);
    if (ctx->ev_film_touch) cudaEventDestroy(ctx->ev_film_touch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);      // (a caller-supplied stream is the caller's to destroy)
    delete ctx;
    return HK_OK;
}
const char* hk_last_error(HkContext* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t hk_upload_tables(HkContext* ctx, const HkTables* t) {
    if (!ctx || !t) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(t->rgb2spec_res >= 2, "rgb2spec_res must be >= 2");
    CK(ctx->b_sobol.upload(t->sobol_matrices, sizeof(uint32_t) * 1024 * 52));
    CK(ctx->b_cie_x.upload(t->cie_x, 4 * 471)); CK(ctx->b_cie_y.upload(t->cie_y, 4 * 471)); CK(ctx->b_cie_z.upload(t->cie_z, 4 * 471));
    CK(ctx->b_d65.upload(t->d65, 4 * 107));
    const size_t R = (size_t)t->rgb2spec_res;
    CK(ctx->b_rgb_scale.upload(t->rgb2spec_scale, 4 * R));
    // re-layout [coef][x][y][z][maxc] -> [maxc][z][y][x] float4(c0,c1,c2,0): see hk_spectral.cuh
    std::vector<float> re(4 * 3 * R * R * R);
    for (size_t m = 0; m < 3; m++) for (size_t z = 0; z < R; z++) for (size_t y = 0; y < R; y++) for (size_t x = 0; x < R; x++) {
        size_t dst = 4 * (((m * R + z) * R + y) * R + x);
        for (size_t c = 0; c < 3; c++) re[dst + c] = t->rgb2spec_coeffs[m + 3 * (z + R * (y + R * (x + R * c)))];
        re[dst + 3] = 0.0f;
    }
    CK(ctx->b_rgb_coeffs.upload(re.data(), re.size() * 4));
    DevTables& T = ctx->D.T;
    T.sobol = ctx->b_sobol.as<uint32_t>(); T.cie_x = ctx->b_cie_x.as<float>(); T.cie_y = ctx->b_cie_y.as<float>(); T.cie_z = ctx->b_cie_z.as<float>();
    T.d65 = ctx->b_d65.as<float>(); T.rgb_scale = ctx->b_rgb_scale.as<float>(); T.rgb_coeffs = ctx->b_rgb_coeffs.as<float4>(); T.rgb_res = t->rgb2spec_res;
    ctx->D.sobol.M = T.sobol;
    {   // closed forms for Sobol' dimensions 0 and 1 are only used when the uploaded table really has that structure
        const uint32_t* M = t->sobol_matrices;
        auto pascal_col = [](int j) { uint32_t r = 0; for (int i = 0; i <= j && i < 32; i++) if ((j & i) == i) r |= 0x80000000u >> i; return r; };   // C(j,i) odd <=> i subset of j
        bool ok = true;
        for (int j = 0; j < 52 && ok; j++) {
            ok = M[j] == (j < 32 ? 0x80000000u >> j : 0u);
            if (ok) ok = M[52 + j] == pascal_col(j & 31) && (j < 32 || j - 32 < 32);
        }
        ctx->D.sobol.fast = ok ? 1 : 0;
    }
    ctx->have_tables = true;
    return hk_refresh_uplift_cache(ctx);
}

// (Re)build the uplift cache for whatever is uploaded: needs the rgb2spec table; called after every upload that changes
// a constant colour (tables, materials, lights, media, hk_update_material).
int32_t hk_refresh_uplift_cache(HkContext* ctx) {
    DevTables& T = ctx->D.T;
    T.mat_pre = nullptr; T.light_pre = nullptr; T.med_pre = nullptr; T.mat_base = nullptr; T.light_base = nullptr; T.med_base = nullptr;
    if (!ctx->have_tables || !ctx->uplift_cache_enabled) return HK_OK;
    const uint32_t nm = ctx->have_mats ? (uint32_t)ctx->mat_types.size() : 0u;
    const uint32_t nl = ctx->have_lights ? (uint32_t)ctx->D.n_lights : 0u;
    const uint32_t nd = ctx->D.media ? (uint32_t)ctx->D.n_media : 0u;
    if (nm + nl + nd == 0) return HK_OK;
    CK(ctx->b_mat_pre.alloc(32 * (size_t)nm)); CK(ctx->b_light_pre.alloc(32 * (size_t)nl)); CK(ctx->b_med_pre.alloc(48 * (size_t)nd));
    k_precompute_uplifts<<<grid_for(ctx, nm + nl + nd, 128, 8), 128, 0, ctx->stream>>>(T, ctx->D.materials, nm, ctx->b_mat_pre.as<float4>(), ctx->D.lights, nl,
                                                                                ctx->b_light_pre.as<float4>(), ctx->D.media, nd, ctx->b_med_pre.as<float4>());
    ctx->launches++;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (nm) { T.mat_pre = ctx->b_mat_pre.as<float4>(); T.mat_base = ctx->D.materials; }
    if (nl) { T.light_pre = ctx->b_light_pre.as<float4>(); T.light_base = ctx->D.lights; }
    if (nd) { T.med_pre = ctx->b_med_pre.as<float4>(); T.med_base = ctx->D.media; }
    return HK_OK;
}

// material type per BVH triangle (HitRec): needs geometry and materials, runs after whichever is uploaded second
static int32_t patch_tri_types(HkContext* ctx) {
    ctx->tri_types_valid = false;
    if (!ctx->have_geom || !ctx->have_mats) return HK_OK;
    if (ctx->max_iface_in_geom > ctx->n_interfaces) return HK_OK;      // geometry and materials of different scenes: wait for the matching upload
    const uint32_t n = (uint32_t)(ctx->b_tris.bytes / sizeof(HkBvhTri));
    ctx->tri_types_valid = true;
    if (ctx->D.n_inst > 0) {      // instanced: the shading class rides in the instance leaf records
        k_patch_inst_types<<<grid_for(ctx, (size_t)ctx->D.n_inst, 256, 8), 256, 0, ctx->stream>>>(ctx->b_inst_recs.as<float4>(), (uint32_t)ctx->D.n_inst, ctx->D.instances, ctx->D.interfaces, ctx->D.materials);
        ctx->launches++;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return HK_OK;
    }
    if (n == 0 || !ctx->D.tri_meta) return HK_OK;
    k_patch_tri_types<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->b_tris.as<float4>(), n, ctx->D.tri_meta, ctx->D.interfaces, ctx->D.materials);
    ctx->launches++;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return HK_OK;
}

int32_t hk_upload_geometry(HkContext* ctx, const HkGeometry* g) {
    if (!ctx || !g) return HK_ERR_INVALID;
    hk_enter(ctx);
    const bool instanced = g->n_instances > 0;
    REQUIRE(g->n_tris == 0 || (g->positions && g->indices && (instanced || g->tri_meta)), "geometry arrays missing");
    REQUIRE(!instanced || (g->meshes && g->n_meshes > 0 && g->instances), "instanced geometry needs meshes and instances");
    for (size_t i = 0; i < 3 * (size_t)g->n_tris; i++) REQUIRE(g->indices[i] < g->n_verts, "vertex index out of range");
    ctx->max_iface_in_geom = 0;
    uint64_t n_world = g->n_tris;
    if (instanced) {
        n_world = 0;
        for (uint32_t m = 0; m < g->n_meshes; m++) REQUIRE((uint64_t)g->meshes[m].first_tri + g->meshes[m].n_tris <= g->n_tris, "mesh range outside the index array");
        for (uint32_t i = 0; i < g->n_instances; i++) {
            REQUIRE(g->instances[i].mesh < g->n_meshes, "instance references a missing mesh");
            REQUIRE(g->instances[i].medium_interface_idx >= 1, "HkInstance.medium_interface_idx is 1-based");
            ctx->max_iface_in_geom = std::max(ctx->max_iface_in_geom, g->instances[i].medium_interface_idx);
            n_world += g->meshes[g->instances[i].mesh].n_tris;
        }
    } else {
        for (uint32_t i = 0; i < g->n_tris; i++) {
            REQUIRE(g->tri_meta[3 * (size_t)i] >= 1, "TriangleMeta.medium_interface_idx is 1-based");
            ctx->max_iface_in_geom = std::max(ctx->max_iface_in_geom, g->tri_meta[3 * (size_t)i]);
        }
    }
    REQUIRE(n_world < HK_PRIM_MASK, "at most 2^28 - 2 triangles (the hit record keeps the material type in the top 4 bits)");
    HkBvh bvh;
    int depth = 0;
    std::vector<float4> inst_recs; std::vector<DevInstance> dinst; std::vector<uint32_t> ibase;
    if (instanced) {
        std::vector<HkMeshRange> mr(g->n_meshes); std::vector<HkInstanceXf> ix(g->n_instances); std::vector<uint32_t> mesh_root;
        for (uint32_t m = 0; m < g->n_meshes; m++) mr[m] = HkMeshRange{g->meshes[m].first_tri, g->meshes[m].n_tris};
        for (uint32_t i = 0; i < g->n_instances; i++) ix[i] = HkInstanceXf{g->instances[i].mesh, g->instances[i].object_to_world};
        depth = hk_build_scene_bvh(g->positions, g->indices, mr.data(), g->n_meshes, ix.data(), g->n_instances, bvh, mesh_root, nullptr, nullptr) + 1;
        dinst.resize(g->n_instances); ibase.resize(g->n_instances);
        uint32_t base = 0;
        for (uint32_t i = 0; i < g->n_instances; i++) {
            const HkInstance& I = g->instances[i]; DevInstance& d = dinst[i];
            std::memcpy(d.o2w, I.object_to_world, 48); std::memcpy(d.w2o, I.world_to_object, 48);
            d.first_tri = g->meshes[I.mesh].first_tri; d.n_tris = g->meshes[I.mesh].n_tris; d.prim_base = base; d.iface = I.medium_interface_idx;
            ibase[i] = base; base += d.n_tris;
        }
        // leaf records of the top level, in leaf order: tris[k].prim = instance index for k < n_instances
        inst_recs.resize(4 * (size_t)g->n_instances);
        auto u2f = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
        for (uint32_t k = 0; k < g->n_instances; k++) {
            const uint32_t i = bvh.tris[k].prim; const HkInstance& I = g->instances[i];
            for (int r = 0; r < 3; r++) inst_recs[4 * (size_t)k + r] = make_float4(I.world_to_object[4 * r], I.world_to_object[4 * r + 1], I.world_to_object[4 * r + 2], I.world_to_object[4 * r + 3]);
            inst_recs[4 * (size_t)k + 3] = make_float4(u2f(mesh_root[I.mesh]), u2f(dinst[i].prim_base), u2f(0u), u2f(i));
        }
    } else depth = hk_build_bvh8(g->positions, g->indices, g->n_tris, bvh);
    if (depth > HK_SM_STACK + HK_LM_STACK - 1) { ctx->err = "the BVH is deeper than the traversal stack (degenerate geometry: thousands of coincident triangle centroids?)"; return HK_ERR_UNSUPPORTED; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->b_nodes.upload(bvh.nodes.data(), bvh.nodes.size() * sizeof(HkBvhNode)));
    CK(ctx->b_tris.upload(bvh.tris.data(), bvh.tris.size() * sizeof(HkBvhTri)));
    CK(ctx->b_pos.upload(g->positions, 12 * (size_t)g->n_verts));
    if (g->normals) CK(ctx->b_nrm.upload(g->normals, 12 * (size_t)g->n_verts)); else ctx->b_nrm.release();
    if (g->uvs) CK(ctx->b_uvs.upload(g->uvs, 8 * (size_t)g->n_verts)); else ctx->b_uvs.release();
    ctx->D.uvs = g->uvs ? ctx->b_uvs.as<float>() : nullptr;
    CK(ctx->b_idx.upload(g->indices, 12 * (size_t)g->n_tris));
    if (!instanced) CK(ctx->b_meta.upload(g->tri_meta, 12 * (size_t)g->n_tris)); else ctx->b_meta.release();
    ctx->D.bvh.nodes = ctx->b_nodes.as<float4>(); ctx->D.bvh.tris = ctx->b_tris.as<float4>(); ctx->D.bvh.one_bits = 0x3F800000u; ctx->D.bvh.inst = nullptr;
    ctx->D.instances = nullptr; ctx->D.inst_prim_base = nullptr; ctx->D.n_inst = 0;
    if (instanced) {
        CK(ctx->b_inst_recs.upload(inst_recs.data(), inst_recs.size() * sizeof(float4)));
        CK(ctx->b_instances.upload(dinst.data(), dinst.size() * sizeof(DevInstance)));
        CK(ctx->b_inst_base.upload(ibase.data(), ibase.size() * 4));
        ctx->D.bvh.inst = ctx->b_inst_recs.as<float4>();
        ctx->D.instances = ctx->b_instances.as<DevInstance>(); ctx->D.inst_prim_base = ctx->b_inst_base.as<uint32_t>(); ctx->D.n_inst = (int32_t)g->n_instances;
    } else { ctx->b_inst_recs.release(); ctx->b_instances.release(); ctx->b_inst_base.release(); }
    ctx->D.positions = ctx->b_pos.as<float>(); ctx->D.normals = g->normals ? ctx->b_nrm.as<float>() : nullptr;
    ctx->D.indices = ctx->b_idx.as<uint32_t>(); ctx->D.tri_meta = instanced ? nullptr : ctx->b_meta.as<uint32_t>();
    ctx->stats.bvh_nodes = bvh.nodes.size();
    ctx->stats.bvh_bytes = bvh.nodes.size() * sizeof(HkBvhNode) + bvh.tris.size() * sizeof(HkBvhTri) + inst_recs.size() * sizeof(float4);
    ctx->n_world_tris = n_world;
    ctx->have_geom = true; ctx->camera_version++;
    return patch_tri_types(ctx);
}

int32_t hk_upload_spectra(HkContext* ctx, const HkSpectra* s) {
    if (!ctx || !s) return HK_ERR_INVALID;
    hk_enter(ctx);
    uint32_t n = s->n_spectra ? s->offsets[s->n_spectra] : 0;
    CK(ctx->b_spec_l.upload(s->lambdas, 4 * (size_t)n)); CK(ctx->b_spec_v.upload(s->values, 4 * (size_t)n));
    CK(ctx->b_spec_o.upload(s->offsets, 4 * ((size_t)s->n_spectra + 1)));
    ctx->D.spec_lambdas = ctx->b_spec_l.as<float>(); ctx->D.spec_values = ctx->b_spec_v.as<float>(); ctx->D.spec_offsets = ctx->b_spec_o.as<uint32_t>();
    return HK_OK;
}

int32_t hk_upload_textures(HkContext* ctx, const HkTexture* t, uint32_t n) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(n == 0 || t, "textures missing");
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& b : ctx->tex_bufs) b.release();
    ctx->tex_bufs.clear(); ctx->tex_bufs.resize(n);
    std::vector<HkTexture> dev(n);
    bool any_alpha = false;
    for (uint32_t i = 0; i < n; i++) {
        REQUIRE(t[i].rgb && t[i].h >= 1 && t[i].w >= 1, "texture needs data and a positive size");
        const size_t texels = (size_t)t[i].h * t[i].w;
        if (t[i].alpha) {      // colours and the alpha plane share one allocation
            std::vector<float> packed(4 * texels);
            std::memcpy(packed.data(), t[i].rgb, 12 * texels); std::memcpy(packed.data() + 3 * texels, t[i].alpha, 4 * texels);
            CK(ctx->tex_bufs[i].upload(packed.data(), 16 * texels));
            any_alpha = true;
        } else CK(ctx->tex_bufs[i].upload(t[i].rgb, 12 * texels));
        dev[i].rgb = ctx->tex_bufs[i].as<float>(); dev[i].h = t[i].h; dev[i].w = t[i].w;
        dev[i].alpha = t[i].alpha ? ctx->tex_bufs[i].as<float>() + 3 * texels : nullptr;
    }
    ctx->D.has_alpha = any_alpha ? 1 : 0;
    CK(ctx->b_textures.upload(dev.data(), sizeof(HkTexture) * (size_t)n));
    ctx->D.textures = n ? ctx->b_textures.as<HkTexture>() : nullptr; ctx->D.n_textures = (int32_t)n;
    return HK_OK;
}
static int32_t mat_textures_ok(HkContext* ctx, const HkMaterial& m) {
    REQUIRE(!(m.flags & HK_MATFLAG_VERTEX_COLORS) || m.type == HK_MAT_MIX || (m.type == HK_MAT_MATTE && m.tex[0] > 0), "VertexColorTexture is supported for MatteMaterial.Kd only");
    REQUIRE(m.tex[3] == 0, "HkMaterial.tex[3] is unused");
    bool any = false;
    for (int k = 0; k < 3; k++) if (m.tex[k] != 0) { any = true; REQUIRE(m.tex[k] >= 1 && m.tex[k] <= ctx->D.n_textures, "material references a texture that has not been uploaded (hk_upload_textures first)"); }
    for (int k = 0; k < 8; k++) if (m.ftex[k] != 0) { any = true; REQUIRE(m.ftex[k] >= 1 && m.ftex[k] <= ctx->D.n_textures, "material references a texture that has not been uploaded (hk_upload_textures first)"); }
    if (m.type == HK_MAT_MIX) for (int k = 0; k < 3; k++) REQUIRE(m.tex[k] == 0, "a MixMaterial has no RGB parameters to texture");      // (ftex[0] = amount)
    REQUIRE(!((m.flags & HK_MATFLAG_SPECTRAL_ETA_K) && (m.tex[0] != 0 || m.tex[1] != 0)), "eta / k are piecewise-linear spectra: they cannot be textured as well");
    return HK_OK;
}
// textured parameters other than Matte.Kd (that one is the HK_SHADE_MATTE_TEX class): host mirror of material_has_textures()
static bool mat_has_param_textures(const HkMaterial& m) {
    if (m.type == HK_MAT_MIX) return false;
    bool any = (m.tex[0] > 0 && m.type != HK_MAT_MATTE) || m.tex[1] > 0 || m.tex[2] > 0;
    for (int k = 0; k < 8; k++) any = any || m.ftex[k] > 0;
    return any;
}
static uint32_t host_shade_class(const HkMaterial& m);
static void refresh_tex_classes(HkContext* ctx, const HkMaterial* m, uint32_t nm) {
    (void)m; (void)nm;
    uint32_t bits = 0;
    for (size_t i = 0; i < ctx->mat_textured.size(); i++) if (ctx->mat_textured[i] && ctx->mat_types[i] != HK_MAT_MIX) bits |= 1u << ctx->mat_types[i];
    ctx->D.tex_classes = bits;
}
static uint32_t host_shade_class(const HkMaterial& m) { return (m.type == HK_MAT_MATTE && m.tex[0] > 0) ? (uint32_t)HK_SHADE_MATTE_TEX : (uint32_t)m.type; }
static bool mat_type_supported(int32_t t) { return (t >= 1 && t < HK_MAX_MAT_TYPES) || t == HK_MAT_MIX || t == HK_MAT_COATED_CONDUCTOR || t == HK_MAT_COATED_DIFFUSE_TRANSMISSION; }
int32_t hk_upload_materials(HkContext* ctx, const HkMaterial* m, uint32_t nm, const HkMediumInterface* mi, uint32_t ni) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(nm == 0 || m, "materials missing"); REQUIRE(ni == 0 || mi, "interfaces missing");
    uint32_t present = 0; int32_t trans = 0;
    for (uint32_t i = 0; i < nm; i++) {
        REQUIRE(mat_type_supported(m[i].type), "unsupported material type");
        { int32_t rc = mat_textures_ok(ctx, m[i]); if (rc != HK_OK) return rc; }
        if (m[i].type == HK_MAT_MIX) { REQUIRE(m[i].ival[0] >= 1 && (uint32_t)m[i].ival[0] <= nm && m[i].ival[1] >= 1 && (uint32_t)m[i].ival[1] <= nm, "MixMaterial references a missing material"); }
        else present |= 1u << host_shade_class(m[i]);
    }
    for (uint32_t i = 0; i < ni; i++) {
        REQUIRE(mi[i].material >= 1 && mi[i].material <= nm, "interface references a missing material");
        if (mi[i].inside != mi[i].outside) trans = 1;
    }
    CK(ctx->b_mats.upload(m, sizeof(HkMaterial) * (size_t)nm)); CK(ctx->b_ifaces.upload(mi, sizeof(HkMediumInterface) * (size_t)ni));
    ctx->D.materials = ctx->b_mats.as<HkMaterial>(); ctx->D.interfaces = ctx->b_ifaces.as<HkMediumInterface>();
    ctx->D.any_medium_transition = trans; ctx->mat_types_present = present; ctx->n_interfaces = ni;
    ctx->mat_textured.assign(nm, 0);
    for (uint32_t i = 0; i < nm; i++) ctx->mat_textured[i] = mat_has_param_textures(m[i]) ? 1 : 0;
    ctx->mat_types.resize(nm); for (uint32_t i = 0; i < nm; i++) ctx->mat_types[i] = m[i].type == HK_MAT_MIX ? HK_MAT_MIX : (int32_t)host_shade_class(m[i]);
    refresh_tex_classes(ctx, m, nm);
    if (!ctx->b_spec_o.p) { uint32_t zero = 0; CK(ctx->b_spec_o.upload(&zero, 4)); ctx->D.spec_offsets = ctx->b_spec_o.as<uint32_t>(); }
    ctx->have_mats = true; ctx->camera_version++;
    int32_t rc = hk_refresh_uplift_cache(ctx);
    return rc != HK_OK ? rc : patch_tri_types(ctx);
}

// update_material!(scene, idx, new_material), src/scene.jl:109-112: replace ONE material in place (RayMakie's interactive
// mode edits materials between frames); no re-upload of the scene.  index is 1-based into the uploaded material array.
int32_t hk_update_material(HkContext* ctx, uint32_t index, const HkMaterial* m) {
    if (!ctx || !m) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_mats, "hk_upload_materials has not been called");
    REQUIRE(index >= 1 && index <= ctx->mat_types.size(), "material index out of range");
    REQUIRE(mat_type_supported(m->type), "unsupported material type");
    { int32_t rc = mat_textures_ok(ctx, *m); if (rc != HK_OK) return rc; }
    if (m->type == HK_MAT_MIX) REQUIRE(m->ival[0] >= 1 && (size_t)m->ival[0] <= ctx->mat_types.size() && m->ival[1] >= 1 && (size_t)m->ival[1] <= ctx->mat_types.size(), "MixMaterial references a missing material");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->b_mats.as<HkMaterial>() + (index - 1), m, sizeof(HkMaterial), cudaMemcpyHostToDevice));
    const int32_t cls = m->type == HK_MAT_MIX ? HK_MAT_MIX : (int32_t)host_shade_class(*m);      // mat_types holds shading classes
    const bool type_changed = ctx->mat_types[index - 1] != cls;
    ctx->mat_types[index - 1] = cls;
    ctx->mat_textured[index - 1] = mat_has_param_textures(*m) ? 1 : 0;
    refresh_tex_classes(ctx, nullptr, 0);
    { int32_t rc = hk_refresh_uplift_cache(ctx); if (rc != HK_OK) return rc; }
    if (type_changed) {
        uint32_t present = 0;
        for (int32_t t : ctx->mat_types) if (t != HK_MAT_MIX) present |= 1u << t;
        ctx->mat_types_present = present;
        return patch_tri_types(ctx);      // the material type rides in the BVH triangle records
    }
    return HK_OK;
}

int32_t hk_upload_envmaps(HkContext* ctx, const HkEnvMap* maps, uint32_t n) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    for (auto& b : ctx->env_bufs) b.release();
    ctx->env_bufs.clear(); ctx->env_bufs.resize(6 * (size_t)n);
    std::vector<DevEnvMap> dev(n);
    for (uint32_t i = 0; i < n; i++) {
        const HkEnvMap& E = maps[i];
        size_t w = E.w, h = E.h, nu = E.nu, nv = E.nv;
        DevBuf* B = &ctx->env_bufs[6 * (size_t)i];
        CK(B[0].upload(E.rgb, 12 * w * h)); CK(B[1].upload(E.conditional_func, 4 * nu * nv)); CK(B[2].upload(E.conditional_cdf, 4 * (nu + 1) * nv));
        CK(B[3].upload(E.conditional_func_int, 4 * nv)); CK(B[4].upload(E.marginal_func, 4 * nv)); CK(B[5].upload(E.marginal_cdf, 4 * (nv + 1)));
        DevEnvMap& d = dev[i];
        d.rgb = B[0].as<float>(); d.w = E.w; d.h = E.h; std::memcpy(d.rot, E.rotation, 36); std::memcpy(d.scale_rgb, E.scale_rgb, 12);
        d.cfunc = B[1].as<float>(); d.ccdf = B[2].as<float>(); d.cfint = B[3].as<float>(); d.mfunc = B[4].as<float>(); d.mcdf = B[5].as<float>();
        d.mfint = E.marginal_func_int; d.nu = E.nu; d.nv = E.nv;
    }
    CK(ctx->b_env.upload(dev.data(), sizeof(DevEnvMap) * (size_t)n));
    ctx->D.envmaps = ctx->b_env.as<DevEnvMap>();
    return HK_OK;
}

int32_t hk_upload_lights(HkContext* ctx, const HkLight* l, uint32_t n, const HkLightSampler* sm) {
    if (!ctx || !sm) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(n == 0 || l, "lights missing");
    for (uint32_t i = 0; i < n; i++) REQUIRE(l[i].type >= 1 && l[i].type <= 7, "unknown light type");
    CK(ctx->b_lights.upload(l, sizeof(HkLight) * (size_t)n));
    {   // the reference's nodes go up as they are and are converted to the device form (DevLNode) by a kernel
        DevBuf raw; CK(raw.upload(sm->nodes, sizeof(HkLightBVHNode) * (size_t)sm->n_nodes));
        CK(ctx->b_lnodes.alloc(sizeof(DevLNode) * (size_t)sm->n_nodes));
        if (sm->n_nodes) { k_prepare_lnodes<<<grid_for(ctx, sm->n_nodes, 256, 8), 256, 0, ctx->stream>>>(raw.as<HkLightBVHNode>(), sm->n_nodes, ctx->b_lnodes.as<DevLNode>()); ctx->launches++; }
        cudaError_t e = cudaStreamSynchronize(ctx->stream); raw.release(); CK(e); CK(cudaGetLastError());
    }
    CK(ctx->b_trails.upload(sm->light_to_bit_trail, 4 * (size_t)n));
    CK(ctx->b_inf.upload(sm->infinite_light_indices, 4 * (size_t)sm->n_infinite));
    ctx->D.lights = ctx->b_lights.as<HkLight>(); ctx->D.n_lights = (int32_t)n;
    ctx->D.lnodes = ctx->b_lnodes.as<DevLNode>(); ctx->D.bit_trails = ctx->b_trails.as<uint32_t>(); ctx->D.inf_idx = ctx->b_inf.as<int32_t>();
    ctx->D.n_infinite = (int32_t)sm->n_infinite; ctx->D.n_bvh = (int32_t)sm->n_bvh_lights;
    {   // large light sets: light selection + sampling run as their own kernel (k_hit_lights) with the cooperative BVH descent
        const char* e = std::getenv("HK_SPLIT_MIN_LIGHTS");      // development override of the threshold
        ctx->D.split_lights = (int32_t)sm->n_bvh_lights >= (e ? std::atoi(e) : HK_COOP_MIN_LIGHTS) ? 1 : 0;
    }
    std::vector<int32_t> esc;
    for (uint32_t i = 0; i < n; i++) if (l[i].type == HK_LIGHT_ENVIRONMENT || l[i].type == HK_LIGHT_AMBIENT) esc.push_back((int32_t)i);
    CK(ctx->b_esc.upload(esc.data(), 4 * esc.size()));
    ctx->D.esc_idx = ctx->b_esc.as<int32_t>(); ctx->D.n_esc = (int32_t)esc.size();
    ctx->have_lights = true;
    return hk_refresh_uplift_cache(ctx);
}

// one medium: voxel buffers, majorant grid (uploaded, or built on the device when HkMedium.majorant is NULL: k_build_majorant) and
// the empty-cell mask of the majorant grid (k_majorant_mask, one bit per cell)
static int32_t upload_one_medium(HkContext* ctx, uint32_t i, const HkMedium& M, DevMedium& d) {
    REQUIRE(M.type >= 1 && M.type <= 4, "unknown medium type");
    std::memset(&d, 0, sizeof(d));
    d.type = M.type; std::memcpy(d.sigma_a, M.sigma_a_rgb, 12); std::memcpy(d.sigma_s, M.sigma_s_rgb, 12); std::memcpy(d.Le, M.Le_rgb, 12); d.g = M.g;
    std::memcpy(d.bmin, M.bounds_min, 12); std::memcpy(d.bmax, M.bounds_max, 12); std::memcpy(d.medium_from_render, M.medium_from_render, 48);
    DevBuf* B = &ctx->media_bufs[3 * (size_t)i];
    if (M.type == HK_MEDIUM_GRID) {
        REQUIRE(M.density && M.density_res[0] >= 1 && M.density_res[1] >= 1 && M.density_res[2] >= 1, "GridMedium needs a density grid");
        size_t cnt = (size_t)M.density_res[0] * M.density_res[1] * M.density_res[2];
        CK(B[0].upload(M.density, 4 * cnt)); d.density = B[0].as<float>(); std::memcpy(d.dres, M.density_res, 12);
    }
    if (M.type == HK_MEDIUM_RGBGRID) {     // the (up to) three RGB grids share one allocation
        REQUIRE(M.rgb_sigma_a || M.rgb_sigma_s, "RGBGridMedium needs at least one of sigma_a / sigma_s grids (media.jl:1073)");
        REQUIRE(!M.rgb_Le || M.rgb_sigma_a, "RGBGridMedium: Le grid requires a sigma_a grid (media.jl:1075)");
        REQUIRE(M.density_res[0] >= 2 && M.density_res[1] >= 2 && M.density_res[2] >= 2, "RGBGridMedium: grid must be at least 2 voxels per axis");
        const size_t cnt = 3 * (size_t)M.density_res[0] * M.density_res[1] * M.density_res[2];
        const float* src[3] = {M.rgb_sigma_a, M.rgb_sigma_s, M.rgb_Le};
        std::vector<float> packed; size_t off[3];
        for (int k = 0; k < 3; k++) { off[k] = packed.size(); if (src[k]) packed.insert(packed.end(), src[k], src[k] + cnt); }
        CK(B[0].upload(packed.data(), 4 * packed.size()));
        d.rgb_a = src[0] ? B[0].as<float>() + off[0] : nullptr; d.rgb_s = src[1] ? B[0].as<float>() + off[1] : nullptr;
        d.rgb_le = src[2] ? B[0].as<float>() + off[2] : nullptr;
        std::memcpy(d.dres, M.density_res, 12); d.sigma_scale = M.scale; d.le_scale = M.Le_scale;
    }
    int32_t idx_min[3], idx_max[3];      // NanoVDB: the index range of the leaves (majorant build, dense mirror)
    std::memcpy(idx_min, M.nanovdb_index_min, 12); std::memcpy(idx_max, M.nanovdb_index_max, 12);
    if (M.type == HK_MEDIUM_NANOVDB) {
        if (M.nanovdb_buf) {
            REQUIRE(M.nanovdb_bytes > 0, "NanoVDBMedium needs its grid buffer");
            CK(B[2].upload(M.nanovdb_buf, (size_t)M.nanovdb_bytes)); d.nvdb = B[2].as<uint8_t>();
            d.root_off = M.nanovdb_root_offset; d.root_tiles = M.nanovdb_root_tiles;
        } else {      // build_nanovdb_from_dense (nanovdb.jl:602-858) on the device, from the dense volume in `density`
            REQUIRE(M.density && M.density_res[0] >= 1 && M.density_res[1] >= 1 && M.density_res[2] >= 1, "NanoVDBMedium needs its grid buffer, or a dense volume (density, density_res) to build it from");
            const size_t cnt = (size_t)M.density_res[0] * M.density_res[1] * M.density_res[2];
            CK(B[0].upload(M.density, 4 * cnt));
            HkNvdbBuilt nb;
            const int32_t rc = hk_nvdb_build_dense(ctx, B[0].as<float>(), M.density_res, 0.0f, B[2], nb);
            B[0].release();
            if (rc != HK_OK) return rc;
            d.nvdb = B[2].as<uint8_t>(); d.root_off = nb.root_off; d.root_tiles = nb.n_up;
            std::memcpy(idx_min, nb.idx_min, 12); std::memcpy(idx_max, nb.idx_max, 12);
        }
        std::memcpy(d.inv_mat, M.nanovdb_inv_mat, 36); std::memcpy(d.vec, M.nanovdb_vec, 12);
    }
    ctx->dense_bufs[i].release();
    if (M.type == HK_MEDIUM_NANOVDB && !std::getenv("HK_NO_DENSE_MIRROR")) {
        // dense mirror of the tree over [index_min, index_max] + 1 (the +1: the upper trilinear corner of the last voxel), when it fits
        size_t cap = 2ull << 30;
        if (const char* e = std::getenv("HK_DENSE_MIRROR_MAX_MB")) cap = (size_t)std::max(0, atoi(e)) << 20;
        bool ok = true; size_t vox = 1;
        for (int k = 0; k < 3; k++) {
            const long long ext = (long long)idx_max[k] - (long long)idx_min[k] + 2;
            ok = ok && ext >= 2 && ext < (1ll << 20);
            if (ok) { d.dn_min[k] = idx_min[k]; d.dn_ext[k] = (int32_t)ext; vox *= (size_t)ext; ok = vox <= (cap >> 2); }
        }
        if (ok) {
            CK(ctx->dense_bufs[i].alloc(4 * vox));
            d.dense = nullptr;      // (the fill itself walks the tree)
            k_nvdb_densify<<<grid_for(ctx, vox, 256, 16), 256, 0, ctx->stream>>>(d, ctx->dense_bufs[i].as<float>());
            ctx->launches++;
            d.dense = ctx->dense_bufs[i].as<float>();
        } else { d.dense = nullptr; for (int k = 0; k < 3; k++) d.dn_min[k] = d.dn_ext[k] = 0; }
    }
    ctx->mask_bufs[i].release();
    if (M.type == HK_MEDIUM_HOMOGENEOUS) return HK_OK;
    REQUIRE(M.majorant_res[0] >= 1 && M.majorant_res[1] >= 1 && M.majorant_res[2] >= 1, "majorant_res must be at least 1 per axis");
    const size_t cells = (size_t)M.majorant_res[0] * M.majorant_res[1] * M.majorant_res[2];
    REQUIRE(cells < (1ull << 31), "majorant grid too large");
    std::memcpy(d.mres, M.majorant_res, 12);
    if (M.majorant) { CK(B[1].upload(M.majorant, 4 * cells)); d.majorant = B[1].as<float>(); }
    else {      // build_majorant_grid / build_rgb_majorant_grid / build_nanovdb_majorant_grid on the device
        CK(B[1].alloc(4 * cells)); d.majorant = B[1].as<float>();
        MajBuild P; std::memcpy(P.idx_min, idx_min, 12); std::memcpy(P.idx_max, idx_max, 12);
        std::memcpy(P.bmin, M.bounds_min, 12); std::memcpy(P.bmax, M.bounds_max, 12);
        k_build_majorant<<<(unsigned)cells, 128, 0, ctx->stream>>>(d, P, B[1].as<float>());
        ctx->launches++;
    }
    const size_t words = (cells + 31) / 32;
    CK(ctx->mask_bufs[i].alloc(4 * words));
    k_majorant_mask<<<grid_for(ctx, words, 256, 8), 256, 0, ctx->stream>>>(d.majorant, (uint32_t)cells, ctx->mask_bufs[i].as<uint32_t>());
    ctx->launches++;
    d.maj_empty = ctx->mask_bufs[i].as<uint32_t>();
    return HK_OK;
}
// after any change to ctx->media_host: which mask the tracking kernels stage in shared memory (the first one that fits
// HK_SMEM_MASK_WORDS), the device copy of the records, the uplift cache
static int32_t commit_media(HkContext* ctx) {
    std::vector<DevMedium>& dev = ctx->media_host;
    ctx->has_rgbgrid = false; ctx->D.smem_mask_medium = 0; ctx->D.smem_mask_words = 0;
    for (size_t i = 0; i < dev.size(); i++) {
        if (dev[i].type == HK_MEDIUM_RGBGRID) ctx->has_rgbgrid = true;
        if (dev[i].type == HK_MEDIUM_HOMOGENEOUS) continue;
        const size_t words = ((size_t)dev[i].mres[0] * dev[i].mres[1] * dev[i].mres[2] + 31) / 32;
        if (ctx->D.smem_mask_medium == 0 && words <= HK_SMEM_MASK_WORDS && !std::getenv("HK_NO_SMEM_MASK")) { ctx->D.smem_mask_medium = (int32_t)i + 1; ctx->D.smem_mask_words = (uint32_t)words; }
    }
    std::vector<DevMedium> up = dev;
    if (std::getenv("HK_NO_EMPTY_MASK")) { for (auto& d : up) d.maj_empty = nullptr; ctx->D.smem_mask_medium = 0; ctx->D.smem_mask_words = 0; }      // development A/B
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    CK(ctx->b_media.upload(up.data(), sizeof(DevMedium) * up.size()));
    ctx->D.media = ctx->b_media.as<DevMedium>(); ctx->D.n_media = (int32_t)up.size();
    ctx->camera_version++;
    return hk_refresh_uplift_cache(ctx);
}
int32_t hk_upload_media(HkContext* ctx, const HkMedium* m, uint32_t n) {
    if (!ctx || (n > 0 && !m)) return HK_ERR_INVALID;
    hk_enter(ctx);
    for (auto& b : ctx->media_bufs) b.release();
    for (auto& b : ctx->mask_bufs) b.release();
    for (auto& b : ctx->dense_bufs) b.release();
    ctx->media_bufs.clear(); ctx->media_bufs.resize(3 * (size_t)n);
    ctx->mask_bufs.clear(); ctx->mask_bufs.resize(n);
    ctx->dense_bufs.clear(); ctx->dense_bufs.resize(n);
    ctx->media_host.assign(n, DevMedium{});
    ctx->media_ok = false;
    for (uint32_t i = 0; i < n; i++) { int32_t rc = upload_one_medium(ctx, i, m[i], ctx->media_host[i]); if (rc != HK_OK) { ctx->media_host.clear(); ctx->D.n_media = 0; return rc; } }
    const int32_t rc = commit_media(ctx);
    ctx->media_ok = rc == HK_OK;
    return rc;
}
// In-place density update of ONE medium (build_majorant_grid! / build_rgb_majorant_grid!, media.jl:1185-1240, 1498-1530: the host
// swaps the voxel data of a medium and rebuilds its majorant grid; everything else in the scene stays).  `index` is 1-based, the
// record replaces the medium wholesale (same layout as hk_upload_media); its majorant is rebuilt on the device when m->majorant is NULL.
int32_t hk_update_medium(HkContext* ctx, uint32_t index, const HkMedium* m) {
    if (!ctx || !m) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(index >= 1 && index <= ctx->media_host.size(), "medium index out of range");
    ctx->media_ok = false;      // (a failure below may have released buffers the device records still point to)
    int32_t rc = upload_one_medium(ctx, index - 1, *m, ctx->media_host[index - 1]);
    if (rc != HK_OK) return rc;
    rc = commit_media(ctx);
    ctx->media_ok = rc == HK_OK;
    return rc;
}
// the NanoVDB buffer of medium `index` (1-based) as the device holds it (uploaded or device-built): *bytes = its size; copied to `out`
// when out != NULL and capacity suffices
int32_t hk_read_nanovdb(HkContext* ctx, uint32_t index, uint8_t* out, uint64_t capacity, uint64_t* bytes) {
    if (!ctx || !bytes) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(index >= 1 && index <= ctx->media_host.size(), "medium index out of range");
    REQUIRE(ctx->media_host[index - 1].type == HK_MEDIUM_NANOVDB, "not a NanoVDB medium");
    const DevBuf& T = ctx->media_bufs[3 * (size_t)(index - 1) + 2];
    *bytes = T.bytes;
    if (!out) return HK_OK;
    REQUIRE(capacity >= T.bytes, "buffer too small for the NanoVDB tree");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, T.p, T.bytes, cudaMemcpyDeviceToHost));
    return HK_OK;
}
// the majorant grid of medium `index` (1-based) as the device holds it, [rz][ry][rx] (uploaded or device-built)
int32_t hk_read_majorant(HkContext* ctx, uint32_t index, float* out, uint64_t n_cells) {
    if (!ctx || !out) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(index >= 1 && index <= ctx->media_host.size(), "medium index out of range");
    const DevMedium& d = ctx->media_host[index - 1];
    REQUIRE(d.type != HK_MEDIUM_HOMOGENEOUS, "a homogeneous medium has no majorant grid");
    REQUIRE(n_cells == (uint64_t)d.mres[0] * d.mres[1] * d.mres[2], "n_cells does not match the medium's majorant_res");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, d.majorant, 4 * (size_t)n_cells, cudaMemcpyDeviceToHost));
    return HK_OK;
}

int32_t hk_set_camera(HkContext* ctx, const HkCamera* c) {
    if (!ctx || !c) return HK_ERR_INVALID;
    if (!ctx->have_cam || std::memcmp(&ctx->D.camera, c, sizeof(HkCamera)) != 0) ctx->camera_version++;      // (an unchanged camera keeps its detected medium)
    ctx->D.camera = *c; ctx->have_cam = true;
    return HK_OK;
}
int32_t hk_set_filter(HkContext* ctx, const HkFilter* f) {
    if (!ctx || !f) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(f->type >= 1 && f->type <= 5, "unknown filter type");
    DevFilter& F = ctx->D.filter; std::memset(&F, 0, sizeof(F));
    F.type = f->type; F.rx = f->radius[0]; F.ry = f->radius[1]; F.nx = f->nx; F.ny = f->ny;
    if (f->type >= 3) {
        REQUIRE(f->nx > 0 && f->ny > 0 && f->func && f->marginal_cdf && f->marginal_func && f->conditional_cdf, "tabulated filter data missing");
        size_t nx = f->nx, ny = f->ny;
        CK(ctx->b_f_func.upload(f->func, 4 * nx * ny)); CK(ctx->b_f_mcdf.upload(f->marginal_cdf, 4 * (ny + 1)));
        CK(ctx->b_f_mfunc.upload(f->marginal_func, 4 * ny)); CK(ctx->b_f_ccdf.upload(f->conditional_cdf, 4 * ny * (nx + 1)));
        F.func = ctx->b_f_func.as<float>(); F.mcdf = ctx->b_f_mcdf.as<float>(); F.mfunc = ctx->b_f_mfunc.as<float>(); F.ccdf = ctx->b_f_ccdf.as<float>();
        F.dmin_x = f->domain_min[0]; F.dmin_y = f->domain_min[1]; F.dmax_x = f->domain_max[0]; F.dmax_y = f->domain_max[1]; F.func_integral = f->func_integral;
    }
    ctx->have_filter = true;
    return HK_OK;
}

// Path-state pool for n_slots = (samples in flight) x n_pixels.  The film accumulators are separate (alloc_film) so that
// the pool can grow when a later call renders more samples per pass without touching what was accumulated.
static int32_t alloc_film(HkContext* ctx, size_t n_pixels) {
    CK(ctx->b_film.alloc(16 * n_pixels + 64));
    CK(cudaMemset(ctx->b_film.p, 0, ctx->b_film.bytes));
    ctx->S.pixel_rgb = ctx->b_film.as<float>(); ctx->S.pixel_weight = ctx->b_film.as<float>() + 3 * n_pixels;
    for (auto& A : ctx->alts) { A.S.pixel_rgb = ctx->S.pixel_rgb; A.S.pixel_weight = ctx->S.pixel_weight; }      // (the other render lanes accumulate into the same film)
    return HK_OK;
}
static int32_t alloc_state(HkContext* ctx, size_t n_slots) {
    // one slab: 22 float4 arrays, 5 u32/f32 arrays, 16 queues; every array starts 256-byte aligned
    const size_t f4 = 22, w4 = 5, q = 9 + HK_N_HIT_QUEUES;
    size_t rounded = f4 * (((16 * n_slots + 255) / 256) * 256) + (w4 + q) * (((4 * n_slots + 255) / 256) * 256);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->b_state.alloc(rounded));
    char* p = ctx->b_state.as<char>();
    auto take = [&](size_t elt) { char* r = p; p += ((elt * n_slots + 255) / 256) * 256; return r; };
    PathState& S = ctx->S;
    float4** f4s[] = {&S.ray_a, &S.ray_b, &S.hit, &S.lambda, &S.lpdf, &S.beta, &S.r_u, &S.r_l, &S.L, &S.sh_a, &S.sh_b, &S.sh_Ld, &S.sh_ru, &S.sh_rl,
                      &S.med, &S.sh_hit, &S.sh_T, &S.sh_tu, &S.sh_tl, &S.nee_a, &S.nee_b, &S.nee_c};
    for (auto pp : f4s) *pp = reinterpret_cast<float4*>(take(16));
    S.flags = reinterpret_cast<uint32_t*>(take(4)); S.fweight = reinterpret_cast<float*>(take(4)); S.sh_medium = reinterpret_cast<uint32_t*>(take(4));
    S.med_ev = reinterpret_cast<uint32_t*>(take(4)); S.res_mat = reinterpret_cast<uint32_t*>(take(4));
    S.q_ray[0] = reinterpret_cast<uint32_t*>(take(4)); S.q_ray[1] = reinterpret_cast<uint32_t*>(take(4));
    S.q_escaped = reinterpret_cast<uint32_t*>(take(4)); S.q_medium = reinterpret_cast<uint32_t*>(take(4)); S.q_shadow = reinterpret_cast<uint32_t*>(take(4));
    S.q_shadow2 = reinterpret_cast<uint32_t*>(take(4));
    S.q_alpha[0] = reinterpret_cast<uint32_t*>(take(4)); S.q_alpha[1] = reinterpret_cast<uint32_t*>(take(4));
    S.q_lbvh = reinterpret_cast<uint32_t*>(take(4));
    for (int t = 0; t < HK_N_HIT_QUEUES; t++) S.q_hit[t] = reinterpret_cast<uint32_t*>(take(4));
    S.counts = ctx->b_counts.as<uint32_t>();
    S.rays_traced = reinterpret_cast<unsigned long long*>(ctx->b_counts.as<char>() + sizeof(uint32_t) * HK_N_COUNTERS);
    S.path_vertices = S.rays_traced + 1;
    ctx->n_slots = n_slots;
    return HK_OK;
}

// (Re)build the ZSobol prefix cache when resolution / sampler parameters change: one k_sobol_prefix pass over
// pixels x dimensions-in-use, amortised over every sample rendered afterwards (like the BVH build at geometry upload).
// Capped at HK_SOBOL_CACHE_BYTES; bounces beyond the cached depth use the uncached evaluation (same bits).
#ifndef HK_AUTO_SLOTS
#define HK_AUTO_SLOTS (128ull << 20)
#endif
#ifndef HK_SOBOL_CACHE_BYTES
#define HK_SOBOL_CACHE_BYTES (8ull << 30)
#endif
static int32_t build_sobol_cache(HkContext* ctx) {
    const HkRenderParams& P = ctx->params;
    SobolParams& SP = ctx->D.sobol;
    const size_t n_pixels = (size_t)P.width * P.height;
    int32_t depths = 0;
    if (ctx->sobol_cache_enabled && P.sobol_log2_spp >= 0 && P.sobol_n_base4_digits - ((P.sobol_log2_spp + 1) >> 1) <= 16) {
        const size_t max_slots = HK_SOBOL_CACHE_BYTES / (4 * n_pixels);
        depths = max_slots >= 8 ? (int32_t)std::min<size_t>((size_t)P.max_depth, (max_slots - 3) / 5) : 0;
    }
    const int32_t key[6] = {P.width, P.height, P.sobol_log2_spp, P.sobol_n_base4_digits, (int32_t)P.sampler_seed, depths};
    if (std::memcmp(key, ctx->sobol_cache_key, sizeof(key)) == 0 && (depths == 0 || ctx->b_sobol_top.p)) {
        if (depths == 0) { SP.top = nullptr; SP.dimhash = nullptr; SP.n_top = 0; }
        return HK_OK;
    }
    SP.top = nullptr; SP.dimhash = nullptr; SP.n_top = 0; SP.top_stride = (uint32_t)n_pixels;
    std::memcpy(ctx->sobol_cache_key, key, sizeof(key));
    if (depths == 0) { ctx->b_sobol_top.release(); return HK_OK; }
    const int32_t n_slots = 3 + 5 * depths;
    std::vector<int32_t> dims(n_slots);
    dims[0] = 1; dims[1] = 3; dims[2] = 6;                       // volpath.jl:150-170
    static const int off[5] = {1, 3, 4, 6, 7};                   // volpath.jl:253-262
    for (int d = 0; d < depths; d++) for (int j = 0; j < 5; j++) dims[3 + 5 * d + j] = 6 + 7 * d + off[j];
    CK(ctx->b_sobol_dims.upload(dims.data(), dims.size() * 4));
    CK(ctx->b_sobol_dimhash.alloc(16 * (size_t)n_slots));
    CK(ctx->b_sobol_top.alloc(4 * n_pixels * (size_t)n_slots));
    k_sobol_prefix<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->b_sobol_top.as<uint32_t>(), ctx->b_sobol_dimhash.as<uint4>(), ctx->b_sobol_dims.as<int32_t>(), n_slots,
                                                              (uint32_t)n_pixels, P.width, P.sobol_log2_spp, P.sobol_n_base4_digits, P.sampler_seed);
    ctx->launches++;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    SP.top = ctx->b_sobol_top.as<uint32_t>(); SP.dimhash = ctx->b_sobol_dimhash.as<uint4>(); SP.n_top = n_slots;
    return HK_OK;
}

int32_t hk_set_params(HkContext* ctx, const HkRenderParams* p) {
    if (!ctx || !p) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(p->width > 0 && p->height > 0, "width/height must be positive");
    REQUIRE(p->max_depth >= 1 && p->max_depth <= 255, "max_depth must be in [1, 255]");
    ctx->params = *p;
    if (ctx->params.sample_batch < 1) {
        // auto: keep ~HK_AUTO_SLOTS path states in flight.  Deep bounces hold < 1 % of the rays but every stage still costs
        // its latency floor (~0.1 ms: the longest single traversal / shading chain); several samples per pass share it.
        // 128 M slots = 16 samples per pass at 4K (59 GB of path state), up to 64 at 1080p -- sized for 180 GB of HBM3e; the pool
        // is allocated on demand for min(batch, samples requested).  Measured at 4K (B200): 4 -> 16 samples in flight is
        // +8 % on C3, +18 % on C4 (32 bounces, half of the time in bounces that hold < 5 % of the rays), +0 % on C5.
        const size_t auto_b = HK_AUTO_SLOTS / ((size_t)p->width * p->height);
        ctx->params.sample_batch = (int32_t)std::min<size_t>(64, std::max<size_t>(1, auto_b));
    }
    DevScene& D = ctx->D;
    D.width = p->width; D.height = p->height; D.max_depth = p->max_depth; D.regularize = p->regularize; D.max_component_value = p->max_component_value;
    D.sobol.log2_spp = p->sobol_log2_spp; D.sobol.n_base4_digits = p->sobol_n_base4_digits; D.sobol.seed = p->sampler_seed;
    size_t n_pixels = (size_t)p->width * p->height;
    // the pool itself is sized on demand by hk_render_samples* (a caller that renders one sample per call never
    // allocates more than one sample's worth)
    // a call that only changes integrator parameters (max_depth, regularize, clamp, samples in flight) keeps what the film has
    // accumulated, like the reference, whose render! reads the VolPath fields on every call (volpath.jl:445-520)
    int32_t rc = HK_OK;
    if (!ctx->have_params || ctx->b_film.bytes != 16 * n_pixels + 64) {
        rc = alloc_film(ctx, n_pixels);
        if (rc != HK_OK) return rc;
        ctx->aux_pixels = 0;                 // film.albedo / normal / depth belong to the previous film
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->b_state.release(); ctx->n_slots = 0;
        for (auto& A : ctx->alts) { A.b_state.release(); A.n_slots = 0; }
    }
    rc = build_sobol_cache(ctx);
    if (rc != HK_OK) return rc;
    ctx->have_params = true;
    return HK_OK;
}

int32_t hk_clear(HkContext* ctx) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter_film_async(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    CK(cudaMemsetAsync(ctx->b_film.p, 0, ctx->b_film.bytes, ctx->stream));
    if (ctx->b_aux.p) CK(cudaMemsetAsync(ctx->b_aux.p, 0, ctx->b_aux.bytes, ctx->stream));      // clear!(film) also resets albedo / normal / depth
    hk_film_touched(ctx);
    return HK_OK;
}

}  // extern "C"
// The shading kernels of one bounce work on disjoint slots (one queue per material type) and only share the atomic queue
// counters, so they are forked over side streams between the routing and the shadow pass: in deep bounces each of them is a
// latency floor of its own (~30-90 us for a handful of paths) and in early bounces one kernel's tail is filled by the next.
// With stage timers on (hk_set_profiling) they stay on the render stream so that the timings remain attributable.
template <int TYPE> static void launch_shade(HkContext* ctx, const PassArgs& A, int next) {
    if (!(ctx->mat_types_present & (1u << TYPE))) return;
    if (ctx->concurrent_shade && ctx->profiling == 0) {
        const int j = ctx->shade_fork_slot++;
        cudaStream_t s = ctx->shade_streams[j % 3];
        cudaStreamWaitEvent(s, ctx->ev_fork, 0);
        hkl_shade(TYPE, ctx->sm_count * 8, s, ctx->D, ctx->S, A, next, ctx->bounce_par);
        cudaEventRecord(ctx->ev_join[j], s);
        ctx->launches++;
        return;
    }
    StageScope sc(ctx, HK_STAGE_SHADE);
    hkl_shade(TYPE, ctx->sm_count * 8, ctx->stream, ctx->D, ctx->S, A, next, ctx->bounce_par);
}
extern "C" {
int32_t hk_render_samples_strided(HkContext* ctx, int32_t first, int32_t stride, int32_t count) {
    if (!ctx) return HK_ERR_INVALID;
    cudaSetDevice(ctx->device);      // (no host-side join of the render lanes here)
    REQUIRE(ctx->have_tables && ctx->have_geom && ctx->have_mats && ctx->have_lights && ctx->have_cam && ctx->have_filter && ctx->have_params,
            "render called before tables/geometry/materials/lights/camera/filter/params were all uploaded");
    REQUIRE(ctx->tri_types_valid, "geometry references medium interfaces that the uploaded material set does not have");
    REQUIRE(ctx->media_ok, "the last media upload / update failed: upload the media again before rendering");
    REQUIRE(count >= 0 && stride >= 1 && first >= 1, "bad sample range");
    const size_t n_pixels = (size_t)ctx->params.width * ctx->params.height;
    // frame pipelining: one-sample calls alternate between the two render lanes (hk_context.h)
    const bool pipelined = count >= 1 && count <= ctx->lane_max_count && ctx->lanes_ready && ctx->frame_pipeline && ctx->profiling == 0 && ctx->stream == ctx->own_stream;
    const int lane = pipelined ? ctx->next_lane : 0;
    ctx->next_lane = pipelined ? (lane + 1) % HK_N_LANES : 0;
    if (!pipelined) ctx->sync_alt_lanes();      // a batched call: plain single-stream order
    struct LaneGuard { HkContext* c; int k; ~LaneGuard() { if (k > 0) c->swap_lane(k); } } lane_guard{ctx, lane};
    if (lane > 0) ctx->swap_lane(lane);
    ctx->last_lane = lane;
    cudaStream_t st = ctx->stream;
    {
        const size_t need = n_pixels * (size_t)std::max<int32_t>(1, std::min<int32_t>(ctx->params.sample_batch, count));
        if (ctx->n_slots < need) { int32_t rc = alloc_state(ctx, need); if (rc != HK_OK) return rc; }
    }
    // detect_camera_medium (intersection.jl:690-747), hoisted out of the per-sample path (reference: one alloc + host sync per sample,
    // volpath.jl:503): re-run only when the camera or the scene changed, on this lane's stream, and its result stays on the device
    // (k_camera reads it from there): no host round trip
    uint32_t* cam_medium = ctx->b_scratch_u32.as<uint32_t>() + lane;
    if (ctx->lane_cam_version[lane] != ctx->camera_version) {
        hkl_detect_camera_medium(st, ctx->D, cam_medium);
        ctx->launches++;
        ctx->lane_cam_version[lane] = ctx->camera_version;
    }
    const bool opaque_only = !ctx->D.any_medium_transition && ctx->D.n_media == 0 && !ctx->D.has_alpha;      // (an alpha-tested surface can let a shadow ray through: the segment walk)
    const bool cnt = (ctx->profiling & 2) != 0;
    unsigned long long* work = ctx->b_work_ctr.as<unsigned long long>();
    const int tgrid = ctx->sm_count * (ctx->D.bvh.inst ? HK_TRACE_BLOCKS_PER_SM_INST : HK_TRACE_BLOCKS_PER_SM);      // persistent: one resident wave
    CK(cudaEventRecord(ctx->ev0, st));
    int32_t done = 0;
    while (done < count) {
        PassArgs A;
        A.n_batch = std::min<int32_t>(ctx->params.sample_batch, count - done);
        A.first_sample = first + done * stride; A.stride = stride; A.n_pixels = (uint32_t)n_pixels;
        const size_t n_slots = n_pixels * (size_t)A.n_batch;
        { StageScope sc(ctx, HK_STAGE_CAMERA); k_camera<<<grid_for(ctx, n_slots, 256, 8), 256, 0, st>>>(ctx->D, ctx->S, A, cam_medium); }
        int cur = 0;
        // Opaque-only scenes, stage timers off: the shadow pass of bounce b runs on its own stream and overlaps reset / trace /
        // route of bounce b+1 (it only reads the shadow records and adds to L; escaped / shading of b+1, which also add to L
        // and rewrite the shadow records, wait for it).  Its three counters are double-buffered by bounce parity.
        const bool overlap_shadow = opaque_only && ctx->shadow_stream != nullptr && ctx->profiling == 0 && ctx->D.n_lights > 0;
        bool shadow_in_flight = false;
        for (int depth = 0; depth < ctx->params.max_depth; depth++) {
            const int par = overlap_shadow ? (depth & 1) : 0;
            ctx->bounce_par = par;
            k_reset_bounce<<<1, HK_N_COUNTERS, 0, st>>>(ctx->S, cur, overlap_shadow ? (par ^ 1) : -1); ctx->launches++;
            if (depth >= 1 && ctx->sort_rays) { StageScope sc(ctx, HK_STAGE_ROUTE); k_sort_rays<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->S, cur); }      // (primary rays are coherent as generated)
            {
                StageScope sc(ctx, HK_STAGE_TRACE);
                hkl_trace(cnt, tgrid, st, ctx->D, ctx->S, cur, 0, work);
            }
            { StageScope sc(ctx, HK_STAGE_ROUTE); k_route<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->D, ctx->S, cur, par, 0); }
            if (ctx->D.has_alpha) for (int r = 1; r < HK_ALPHA_ROUNDS; r++) {      // alpha-tested surfaces: re-trace the rays whose hit was skipped (empty rounds exit at once)
                { StageScope sc(ctx, HK_STAGE_TRACE); hkl_trace(cnt, tgrid, st, ctx->D, ctx->S, cur, r, work); }
                { StageScope sc(ctx, HK_STAGE_ROUTE); k_route<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->D, ctx->S, cur, par, r); }
            }
            if (ctx->D.n_media > 0) {
                StageScope sc(ctx, HK_STAGE_MEDIUM);
                hkl_medium_track(ctx->has_rgbgrid, ctx->sm_count * 4, st, ctx->D, ctx->S);
                hkl_medium_finish(ctx->sm_count * 8, st, ctx->D, ctx->S, A, cur ^ 1); ctx->launches++;
            }
            if (shadow_in_flight) { cudaStreamWaitEvent(st, ctx->ev_shadowed, 0); shadow_in_flight = false; }      // shadow(b-1) done before anything adds to L
            const bool fork = ctx->concurrent_shade && ctx->profiling == 0;
            if (ctx->D.n_lights > 0 && ctx->D.split_lights) {      // emissive-hit MIS + the light sample of every surface hit of the bounce, ahead of the per-material kernels
                StageScope sc(ctx, HK_STAGE_SHADE);
                hkl_hit_lights(ctx->sm_count * 4, st, ctx->D, ctx->S, A);
                if (HK_LIGHTS_COMPACT) { hkl_hit_lights_bvh(ctx->sm_count * 6, st, ctx->D, ctx->S, A); ctx->launches++; }
            }
            if (fork) { ctx->shade_fork_slot = 0; cudaEventRecord(ctx->ev_fork, st); }      // (the escaped-ray kernel overlaps the shading kernels too)
            if (ctx->D.n_lights > 0) { StageScope sc(ctx, HK_STAGE_ESCAPED); k_escaped<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->D, ctx->S); }
            launch_shade<HK_MAT_MATTE>(ctx, A, cur ^ 1); launch_shade<HK_MAT_MIRROR>(ctx, A, cur ^ 1); launch_shade<HK_MAT_GLASS>(ctx, A, cur ^ 1);
            launch_shade<HK_MAT_CONDUCTOR>(ctx, A, cur ^ 1); launch_shade<HK_MAT_COATED_DIFFUSE>(ctx, A, cur ^ 1);
            launch_shade<HK_MAT_THIN_DIELECTRIC>(ctx, A, cur ^ 1); launch_shade<HK_MAT_DIFFUSE_TRANSMISSION>(ctx, A, cur ^ 1);
            launch_shade<HK_MAT_COATED_CONDUCTOR>(ctx, A, cur ^ 1); launch_shade<HK_MAT_COATED_DIFFUSE_TRANSMISSION>(ctx, A, cur ^ 1);
            launch_shade<HK_SHADE_MATTE_TEX>(ctx, A, cur ^ 1);
            if (fork) for (int j = 0; j < ctx->shade_fork_slot; j++) cudaStreamWaitEvent(st, ctx->ev_join[j], 0);
            if (ctx->D.n_lights > 0) {
                if (overlap_shadow) {
                    cudaEventRecord(ctx->ev_shaded, st);
                    cudaStreamWaitEvent(ctx->shadow_stream, ctx->ev_shaded, 0);
                    hkl_shadow_opaque(false, tgrid, ctx->shadow_stream, ctx->D, ctx->S, work, par);
                    cudaEventRecord(ctx->ev_shadowed, ctx->shadow_stream);
                    ctx->launches++; shadow_in_flight = true;
                } else {
                    StageScope sc(ctx, HK_STAGE_SHADOW);
                    if (opaque_only) hkl_shadow_opaque(cnt, tgrid, st, ctx->D, ctx->S, work, par);
                    else for (int r = 0; r < HK_SHADOW_ROUNDS; r++) {      // one round per medium-boundary crossing; empty rounds exit at once
                        hkl_shadow_seg_trace(cnt, tgrid, st, ctx->D, ctx->S, r, work);
                        hkl_shadow_seg_ratio(ctx->has_rgbgrid, ctx->sm_count * 4, st, ctx->D, ctx->S, r);
                        ctx->launches += r == 0 ? 1 : 2;
                    }
                }
            }
            if ((ctx->profiling & 5) == 5) {   // per-bounce record (debug): counts as left by this bounce + its stage times
                if ((int)ctx->bounce_counts.size() <= depth) { ctx->bounce_counts.resize(depth + 1); ctx->bounce_ms.resize(depth + 1); }
                ctx->bounce_ms[depth].fill(0.0);
                collect_stage_times(ctx, ctx->bounce_ms[depth].data());
                cudaMemcpy(ctx->bounce_counts[depth].data(), ctx->S.counts, sizeof(uint32_t) * HK_N_QUEUE_COUNTERS, cudaMemcpyDeviceToHost);
            }
            cur ^= 1;
        }
        if (shadow_in_flight) cudaStreamWaitEvent(st, ctx->ev_shadowed, 0);      // the film pass reads L
        // the film is summed in sample order: wait for the previous frame's accumulation, and for read-outs / clears of the film (main or copy stream)
        if (ctx->last_accum_lane >= 0 && ctx->last_accum_lane != lane && ctx->lane_pending[ctx->last_accum_lane]) cudaStreamWaitEvent(st, ctx->ev_lane_film[ctx->last_accum_lane], 0);
        if (ctx->film_touch_pending) cudaStreamWaitEvent(st, ctx->ev_film_touch, 0);      // (a no-op when the event was recorded on this stream)
        { StageScope sc(ctx, HK_STAGE_FILM); k_film_accumulate<<<grid_for(ctx, n_pixels, 256, 8), 256, 0, st>>>(ctx->D, ctx->S, A); }
        if (pipelined || ctx->any_alt_pending()) { cudaEventRecord(ctx->ev_lane_film[lane], st); ctx->lane_pending[lane] = true; ctx->last_accum_lane = lane; }
        done += A.n_batch;
        if (ctx->profiling & 1) collect_stage_times(ctx);
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaGetLastError());
    ctx->stats.samples_rendered += (uint64_t)count * n_pixels;
    return HK_OK;
}
// profiling: bit 0 = time every stage launch with CUDA events, bit 1 = count traversal node visits / triangle tests
int32_t hk_set_profiling(HkContext* ctx, int32_t mode) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    ctx->profiling = mode;
    for (int i = 0; i < HK_N_STAGES; i++) { ctx->stage_ms[i] = 0; ctx->stage_launches[i] = 0; }
    CK(cudaMemset(ctx->b_work_ctr.p, 0, 64));
    return HK_OK;
}
int32_t hk_stage_times(HkContext* ctx, double* out_ms, uint64_t* out_launches, uint64_t* out_work) {
    if (!ctx || !out_ms || !out_launches || !out_work) return HK_ERR_INVALID;
    hk_enter(ctx);
    collect_stage_times(ctx);
    for (int i = 0; i < HK_N_STAGES; i++) { out_ms[i] = ctx->stage_ms[i]; out_launches[i] = ctx->stage_launches[i]; }
    CK(cudaMemcpy(out_work, ctx->b_work_ctr.p, 48, cudaMemcpyDeviceToHost));
    return HK_OK;
}
// per-bounce profile of the most recent sample pass rendered with hk_set_profiling(ctx, 5): counts[max_depth][16] (queue
// counters after the bounce: ray0, ray1, escaped, medium, shadow, total hits, cursors, per-material hits) and ms[max_depth][8]
int32_t hk_bounce_profile(HkContext* ctx, int32_t max_depth, uint32_t* counts, double* ms) {
    if (!ctx || !counts || !ms) return HK_ERR_INVALID;
    for (int d = 0; d < max_depth; d++) for (int i = 0; i < HK_N_QUEUE_COUNTERS; i++) counts[d * HK_N_QUEUE_COUNTERS + i] = d < (int)ctx->bounce_counts.size() ? ctx->bounce_counts[d][i] : 0u;
    for (int d = 0; d < max_depth; d++) for (int i = 0; i < HK_N_STAGES; i++) ms[d * HK_N_STAGES + i] = d < (int)ctx->bounce_ms.size() ? ctx->bounce_ms[d][i] : 0.0;
    return HK_OK;
}
int32_t hk_render_samples(HkContext* ctx, int32_t first, int32_t count) { return hk_render_samples_strided(ctx, first, 1, count); }

int32_t hk_synchronize(HkContext* ctx) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return HK_OK;
}

int32_t hk_read_film(HkContext* ctx, float* out) {
    if (!ctx || !out) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    if (ctx->b_readback.bytes < 12 * n) CK(ctx->b_readback.alloc(12 * n));
    k_film_finalize<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, ctx->b_readback.as<float>(), ctx->params.width, ctx->params.height);
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->b_readback.p, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));   // true async DMA when `out` is pinned
    CK(cudaStreamSynchronize(ctx->stream));
    return HK_OK;
}
int32_t hk_read_film_dev(HkContext* ctx, float* out_dev) {
    if (!ctx || !out_dev) return HK_ERR_INVALID;
    hk_enter_film_async(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out_dev) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError(); ctx->err = "hk_read_film_dev needs a device pointer (use hk_read_film for host memory)"; return HK_ERR_INVALID;
    }
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    k_film_finalize<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, out_dev, ctx->params.width, ctx->params.height);
    ctx->launches++;
    hk_film_touched(ctx);
    CK(cudaGetLastError());
    return HK_OK;
}
int32_t hk_set_stream(HkContext* ctx, void* cuda_stream) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return HK_OK;
}
// Pipelined read-out for progressive display: finalize into one of HK_N_READOUTS staging buffers and copy it to the (pinned) host
// buffer, both on the library's copy stream, and return at once -- the next hk_render_samples calls can be enqueued while the DMA
// runs.  The finalize is ordered after everything enqueued on the render stream so far and after the last lane's film accumulation;
// it does NOT sit in the render stream: there the main lane's next frame would queue behind it and wait for every frame in flight
// (with three read-outs in flight the loop is bound by the host's wait for the oldest frame and the two placements measure the
// same, 958-975 Msamples/s on C3; profiles/r02_e2e_lanes.txt).  Later accumulations / film readers wait for ev_film_touch.  hk_read_film_wait(ticket) blocks until that frame has landed.  At most HK_N_READOUTS (4) frames in flight.
int32_t hk_read_film_async(HkContext* ctx, float* out_pinned, int32_t* ticket) {
    if (!ctx || !out_pinned || !ticket) return HK_ERR_INVALID;
    cudaSetDevice(ctx->device);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    REQUIRE(ctx->ev_film_touch != nullptr, "the context has no film events (creation failed)");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < HK_N_READOUTS; i++) { CK(cudaEventCreateWithFlags(&ctx->ev_final[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming)); }
    }
    cudaStream_t fs = ctx->copy_stream;
    const int k = ctx->async_next;
    if (ctx->b_readback_async[k].bytes < 12 * n) { CK(cudaStreamSynchronize(fs)); CK(ctx->b_readback_async[k].alloc(12 * n)); ctx->async_used[k] = false; }
    CK(cudaEventRecord(ctx->ev_final[k], ctx->stream));      // everything on the render stream so far (earlier frames, clears, batched calls) ...
    CK(cudaStreamWaitEvent(fs, ctx->ev_final[k], 0));
    if (ctx->last_accum_lane > 0 && ctx->lane_pending[ctx->last_accum_lane]) CK(cudaStreamWaitEvent(fs, ctx->ev_lane_film[ctx->last_accum_lane], 0));      // ... and the last accumulation
    if (ctx->film_touch_pending) CK(cudaStreamWaitEvent(fs, ctx->ev_film_touch, 0));
    // (the staging buffer's previous copy ran earlier on this same stream)
    k_film_finalize<<<grid_for(ctx, n, 256, 8), 256, 0, fs>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, ctx->b_readback_async[k].as<float>(), ctx->params.width, ctx->params.height);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_film_touch, fs)); ctx->film_touch_pending = true;
    CK(cudaMemcpyAsync(out_pinned, ctx->b_readback_async[k].p, 12 * n, cudaMemcpyDeviceToHost, fs));
    CK(cudaEventRecord(ctx->ev_copied[k], fs));
    ctx->async_used[k] = true; ctx->async_next = (k + 1) % HK_N_READOUTS;
    *ticket = k;
    return HK_OK;
}
int32_t hk_read_film_wait(HkContext* ctx, int32_t ticket) {
    if (!ctx || ticket < 0 || ticket >= HK_N_READOUTS) return HK_ERR_INVALID;
    cudaSetDevice(ctx->device);
    REQUIRE(ctx->copy_stream && ctx->async_used[ticket], "no asynchronous read-out with this ticket is in flight");
    CK(cudaEventSynchronize(ctx->ev_copied[ticket]));
    return HK_OK;
}
int32_t hk_postprocess(HkContext* ctx, const HkPostprocess* p, float* out) {
    if (!ctx || !p || !out) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    REQUIRE(p->tonemap_mode >= HK_TONEMAP_NONE && p->tonemap_mode <= HK_TONEMAP_FILMIC, "unknown tonemap mode");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    if (ctx->b_readback.bytes < 12 * n) CK(ctx->b_readback.alloc(12 * n));
    REQUIRE(!p->mask_escaped || ctx->aux_pixels == n, "postprocess with a background needs film.depth: call hk_fill_aux_buffers first");
    k_film_postprocess<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, ctx->b_readback.as<float>(), ctx->params.width, ctx->params.height, *p,
                                                                           p->mask_escaped ? ctx->b_aux.as<float>() + 6 * n : nullptr);
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->b_readback.p, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HK_OK;
}
int32_t hk_postprocess_dev(HkContext* ctx, const HkPostprocess* p, float* out_dev) {
    if (!ctx || !p || !out_dev) return HK_ERR_INVALID;
    hk_enter_film_async(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    REQUIRE(p->tonemap_mode >= HK_TONEMAP_NONE && p->tonemap_mode <= HK_TONEMAP_FILMIC, "unknown tonemap mode");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    REQUIRE(!p->mask_escaped || ctx->aux_pixels == n, "postprocess with a background needs film.depth: call hk_fill_aux_buffers first");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out_dev) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError(); ctx->err = "hk_postprocess_dev needs a device pointer"; return HK_ERR_INVALID;
    }
    k_film_postprocess<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, out_dev, ctx->params.width, ctx->params.height, *p,
                                                                           p->mask_escaped ? ctx->b_aux.as<float>() + 6 * n : nullptr);
    ctx->launches++;
    hk_film_touched(ctx);
    CK(cudaGetLastError());
    return HK_OK;
}
// fill_aux_buffers!(film, scene, camera; has_infinite_lights), film.jl:410-431
int32_t hk_fill_aux_buffers(HkContext* ctx, int32_t has_infinite_lights) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    REQUIRE(ctx->have_geom && ctx->have_cam, "geometry and camera must be uploaded before hk_fill_aux_buffers");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    if (ctx->b_aux.bytes != 28 * n) { CK(cudaStreamSynchronize(ctx->stream)); CK(ctx->b_aux.alloc(28 * n)); }
    ctx->aux_pixels = n;
    float* a = ctx->b_aux.as<float>();
    hkl_aux_buffers(grid_for(ctx, n, HK_TRACE_THREADS, 8), ctx->stream, ctx->D, a, a + 3 * n, a + 6 * n, has_infinite_lights ? 1.0e30f : std::numeric_limits<float>::infinity());
    ctx->launches++;
    CK(cudaGetLastError());
    return HK_OK;
}
int32_t hk_read_aux_buffers(HkContext* ctx, float* albedo, float* normal, float* depth) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    REQUIRE(ctx->aux_pixels == n, "hk_fill_aux_buffers has not been called for this film size");
    const float* a = ctx->b_aux.as<float>();
    if (albedo) CK(cudaMemcpyAsync(albedo, a, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (normal) CK(cudaMemcpyAsync(normal, a + 3 * n, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (depth) CK(cudaMemcpyAsync(depth, a + 6 * n, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HK_OK;
}
// denoise!(film; config), denoise.jl:301-372
int32_t hk_denoise(HkContext* ctx, const HkDenoiseConfig* cfg, float* out_pp, float* out_fb) {
    if (!ctx || !cfg || !out_pp) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    REQUIRE(cfg->iterations >= 0 && cfg->iterations <= 16, "denoise iterations out of range (0..16)");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    REQUIRE(ctx->aux_pixels == n, "denoise needs film.normal / film.depth: call hk_fill_aux_buffers first");
    const int W = ctx->params.width, H = ctx->params.height;
    if (ctx->b_denoise.bytes != 28 * n) { CK(cudaStreamSynchronize(ctx->stream)); CK(ctx->b_denoise.alloc(28 * n)); }
    float* buf_a = ctx->b_denoise.as<float>();            // film.framebuffer
    float* buf_b = buf_a + 3 * n;                         // similar(film.framebuffer)
    float* var = buf_a + 6 * n;
    const float* aux = ctx->b_aux.as<float>();
    const dim3 grid = grid_for(ctx, n, 256, 8);
    k_film_finalize<<<grid, 256, 0, ctx->stream>>>(ctx->S.pixel_rgb, ctx->S.pixel_weight, buf_a, W, H);
    if (cfg->use_variance) k_denoise_variance<<<grid, 256, 0, ctx->stream>>>(var, buf_a, W, H);
    ctx->launches += cfg->use_variance ? 2 : 1;
    for (int i = 1; i <= cfg->iterations; i++) {
        DenoisePass P{W, H, 1 << (i - 1), cfg->sigma_color, cfg->sigma_normal, cfg->sigma_depth, cfg->use_variance ? 1 : 0};
        if (i % 2 == 1) k_denoise_atrous<<<grid, 256, 0, ctx->stream>>>(buf_b, buf_a, aux + 3 * n, aux + 6 * n, var, P);
        else k_denoise_atrous<<<grid, 256, 0, ctx->stream>>>(buf_a, buf_b, aux + 3 * n, aux + 6 * n, var, P);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_pp, cfg->iterations % 2 == 1 ? buf_b : buf_a, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_fb && cfg->iterations >= 2) CK(cudaMemcpyAsync(out_fb, buf_a, 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HK_OK;
}
// pinned host memory for film.framebuffer (the Julia shim would use CUDA.pin / cudaHostRegister on the Film's array)
int32_t hk_pinned_alloc(uint64_t bytes, void** out) {
    if (!out) return HK_ERR_INVALID;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault) == cudaSuccess ? HK_OK : HK_ERR_CUDA;
}
int32_t hk_pinned_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? HK_OK : HK_ERR_CUDA; }
int32_t hk_film_accum_dev(HkContext* ctx, float** out, uint64_t* count) {
    if (!ctx || !out || !count) return HK_ERR_INVALID;
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    *out = ctx->S.pixel_rgb; *count = 4ull * (uint64_t)ctx->params.width * ctx->params.height;
    return HK_OK;
}
int32_t hk_read_accum(HkContext* ctx, float* rgb, float* w) {
    if (!ctx || !rgb || !w) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(rgb, ctx->S.pixel_rgb, 12 * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(w, ctx->S.pixel_weight, 4 * n, cudaMemcpyDeviceToHost));
    return HK_OK;
}
int32_t hk_write_accum(HkContext* ctx, const float* rgb, const float* w) {
    if (!ctx || !rgb || !w) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_params, "hk_set_params has not been called");
    const size_t n = (size_t)ctx->params.width * ctx->params.height;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->S.pixel_rgb, rgb, 12 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ctx->S.pixel_weight, w, 4 * n, cudaMemcpyHostToDevice));
    return HK_OK;
}

// ---- stand-alone traversal -----------------------------------------------------------------------------------------
static int32_t trace_dev(HkContext* ctx, const float* rays_dev, uint64_t n64, float* hits_dev, uint8_t* occ_dev, int repeat, bool any, bool count) {
    cudaStream_t st = ctx->stream;
    REQUIRE(n64 < 0xFFFFFF00ull, "at most 2^32-256 rays per call");
    const uint32_t n = (uint32_t)n64;
    unsigned long long* ctr = ctx->b_trace_ctr.as<unsigned long long>();
    CK(cudaEventRecord(ctx->ev0, st));
    for (int r = 0; r < repeat; r++) {
        CK(cudaMemsetAsync(ctr, 0, 24, st));
        const int grid = ctx->sm_count * HK_TRACE_BLOCKS_PER_SM;
        const float4* rp = reinterpret_cast<const float4*>(rays_dev); float4* hp = reinterpret_cast<float4*>(hits_dev);
        hkl_trace_batch(any, count, grid, st, ctx->D.bvh, rp, n, hp, occ_dev, reinterpret_cast<uint32_t*>(ctr), ctr + 1);
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.last_trace_ms = ms / (float)(repeat > 0 ? repeat : 1);
    ctx->stats.rays_traced += (uint64_t)n * (uint64_t)repeat;
    return HK_OK;
}
int32_t hk_trace_closest_dev(HkContext* ctx, const float* rays_dev, uint64_t n, float* hits_dev, int32_t repeat) {
    if (!ctx || !rays_dev || !hits_dev) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_geom, "hk_upload_geometry has not been called");
    return trace_dev(ctx, rays_dev, n, hits_dev, nullptr, repeat < 1 ? 1 : repeat, false, false);
}
int32_t hk_trace_closest(HkContext* ctx, const float* rays, uint64_t n, float* hits) {
    if (!ctx || (n && (!rays || !hits))) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_geom, "hk_upload_geometry has not been called");
    if (n == 0) return HK_OK;
    DevBuf r, h; CK(r.upload(rays, 32 * (size_t)n)); CK(h.alloc(16 * (size_t)n));
    int32_t rc = trace_dev(ctx, r.as<float>(), n, h.as<float>(), nullptr, 1, false, false);
    if (rc == HK_OK) { cudaError_t e = cudaMemcpy(hits, h.p, 16 * (size_t)n, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = HK_ERR_CUDA; } }
    r.release(); h.release();
    return rc;
}
int32_t hk_trace_any(HkContext* ctx, const float* rays, uint64_t n, uint8_t* occluded) {
    if (!ctx || (n && (!rays || !occluded))) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_geom, "hk_upload_geometry has not been called");
    if (n == 0) return HK_OK;
    DevBuf r, o; CK(r.upload(rays, 32 * (size_t)n)); CK(o.alloc((size_t)n));
    int32_t rc = trace_dev(ctx, r.as<float>(), n, nullptr, o.as<uint8_t>(), 1, true, false);
    if (rc == HK_OK) { cudaError_t e = cudaMemcpy(occluded, o.p, (size_t)n, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = HK_ERR_CUDA; } }
    r.release(); o.release();
    return rc;
}
// traversal work counters for the roofline formula: node visits and triangle tests of one batch
int32_t hk_test_trace_counts(HkContext* ctx, const float* rays, uint64_t n, uint64_t* out_nodes, uint64_t* out_tris) {
    if (!ctx || !rays || !out_nodes || !out_tris) return HK_ERR_INVALID;
    hk_enter(ctx);
    REQUIRE(ctx->have_geom, "hk_upload_geometry has not been called");
    DevBuf r, h; CK(r.upload(rays, 32 * (size_t)n)); CK(h.alloc(16 * (size_t)n + 16));
    int32_t rc = trace_dev(ctx, r.as<float>(), n, h.as<float>(), nullptr, 1, false, true);
    unsigned long long c[2] = {0, 0};
    // counters live after the cursor word
    if (rc == HK_OK) { cudaMemcpy(c, ctx->b_trace_ctr.as<unsigned long long>() + 1, 16, cudaMemcpyDeviceToHost); }
    *out_nodes = c[0]; *out_tris = c[1];
    r.release(); h.release();
    return rc;
}

int32_t hk_stats(HkContext* ctx, HkStats* out) {
    if (!ctx || !out) return HK_ERR_INVALID;
    hk_enter(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.last_render_ms = ms; else cudaGetLastError();
    if (ctx->last_lane > 0 && ctx->alts[ctx->last_lane - 1].ev0) { float m2 = 0; if (cudaEventElapsedTime(&m2, ctx->alts[ctx->last_lane - 1].ev0, ctx->alts[ctx->last_lane - 1].ev1) == cudaSuccess) ctx->stats.last_render_ms = m2; else cudaGetLastError(); }
    unsigned long long rt = 0, rt2 = 0;
    if (ctx->S.rays_traced) cudaMemcpy(&rt, ctx->S.rays_traced, 8, cudaMemcpyDeviceToHost);
    for (auto& A : ctx->alts) if (A.S.rays_traced) { unsigned long long x = 0; cudaMemcpy(&x, A.S.rays_traced, 8, cudaMemcpyDeviceToHost); rt2 += x; }
    *out = ctx->stats;
    out->rays_traced = ctx->stats.rays_traced + rt + rt2;
    { unsigned long long pv = 0, pv2 = 0; if (ctx->S.path_vertices) cudaMemcpy(&pv, ctx->S.path_vertices, 8, cudaMemcpyDeviceToHost);
      for (auto& A : ctx->alts) if (A.S.path_vertices) { unsigned long long x = 0; cudaMemcpy(&x, A.S.path_vertices, 8, cudaMemcpyDeviceToHost); pv2 += x; }
      out->path_vertices = pv + pv2; }
    out->kernel_launches = ctx->launches;
    out->queue_overflows = 0;
    return HK_OK;
}

int32_t hk_dev_alloc(HkContext* ctx, uint64_t bytes, void** out) { if (!ctx || !out) return HK_ERR_INVALID; hk_enter(ctx); CK(cudaMalloc(out, bytes ? bytes : 16)); return HK_OK; }
int32_t hk_dev_free(HkContext* ctx, void* p) { if (!ctx) return HK_ERR_INVALID; hk_enter(ctx); CK(cudaFree(p)); return HK_OK; }
int32_t hk_dev_upload(HkContext* ctx, void* dst, const void* src, uint64_t bytes) { if (!ctx) return HK_ERR_INVALID; hk_enter(ctx); CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return HK_OK; }
int32_t hk_dev_download(HkContext* ctx, void* dst, const void* src, uint64_t bytes) { if (!ctx) return HK_ERR_INVALID; hk_enter(ctx); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return HK_OK; }

}  // extern "C"

