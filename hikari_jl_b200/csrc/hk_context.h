// hk_context.h — the context behind the opaque HkContext* of include/hikari_cuda.h, shared by the translation units of
// libhikari_cuda.so (hk_api.cu: uploads, render loop, film; hk_testing.cu: the per-stage test entry points).
#pragma once
#include "../../include/hikari_cuda.h"
#include "../../include/hikari_cuda_testing.h"
#include "hk_wavefront.cuh"
#include <algorithm>
#include <array>
#include <cstring>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return HK_ERR_CUDA; } } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) { ctx->err = (msg); return HK_ERR_INVALID; } } while (0)

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    cudaError_t alloc(size_t n) { release(); if (n == 0) n = 16; cudaError_t e = cudaMalloc(&p, n); if (e == cudaSuccess) bytes = n; else p = nullptr; return e; }
    cudaError_t upload(const void* src, size_t n) { cudaError_t e = alloc(n); if (e != cudaSuccess) return e; return n ? cudaMemcpy(p, src, n, cudaMemcpyHostToDevice) : cudaSuccess; }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HkContext {
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;              // the render stream: own_stream, or the caller's (hk_set_stream)
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    DevScene D;
    PathState S;
    HkRenderParams params;
    bool have_tables = false, have_geom = false, have_mats = false, have_lights = false, have_cam = false, have_filter = false, have_params = false;
    uint64_t camera_version = 1, lane_cam_version[8] = {};      // detect_camera_medium result per render lane (device-resident, b_scratch_u32[lane])
    uint32_t mat_types_present = 0;
    uint32_t n_interfaces = 0, max_iface_in_geom = 0; bool tri_types_valid = false;
    std::vector<int32_t> mat_types;                  // host copy of the material types (hk_update_material)
    std::vector<uint8_t> mat_textured;               // per material: has textured parameters other than Matte.Kd (DevScene::tex_classes)
    // device buffers
    DevBuf b_sobol, b_cie_x, b_cie_y, b_cie_z, b_d65, b_rgb_scale, b_rgb_coeffs;
    DevBuf b_nodes, b_tris, b_pos, b_nrm, b_idx, b_meta;
    DevBuf b_inst_recs, b_instances, b_inst_base; uint64_t n_world_tris = 0;      // instancing (HkGeometry.instances)
    DevBuf b_mats, b_ifaces, b_spec_l, b_spec_v, b_spec_o;
    DevBuf b_lights, b_env, b_lnodes, b_trails, b_inf, b_esc;
    DevBuf b_mat_pre, b_light_pre, b_med_pre;       // uplift cache (DevTables)
    bool uplift_cache_enabled = true;
    std::vector<DevBuf> env_bufs, media_bufs, mask_bufs, dense_bufs;
    bool media_ok = true;                            // false while / after a media upload that failed half-way: render refuses until media are uploaded again
    DevBuf b_media; std::vector<DevMedium> media_host;      // host copy of the device records (hk_update_medium)
    DevBuf b_f_func, b_f_mcdf, b_f_mfunc, b_f_ccdf;
    DevBuf b_state, b_counts, b_rays, b_film, b_scratch_u32, b_trace_ctr, b_readback;
    DevBuf b_aux, b_denoise; size_t aux_pixels = 0;
    DevBuf b_uvs, b_textures; std::vector<DevBuf> tex_bufs;
    bool has_rgbgrid = false;                // some uploaded medium is an RGBGridMedium: the tracking kernels with that branch compiled in     // film.albedo [3n] | film.normal [3n] | film.depth [n], (H, W) column-major
    // pipelined read-out (hk_read_film_async): two device staging buffers, a copy stream, per-buffer events
#ifndef HK_N_READOUTS
#define HK_N_READOUTS 4                // asynchronous read-outs in flight (hk_read_film_async tickets 0..3)
#endif
    DevBuf b_readback_async[HK_N_READOUTS]; cudaStream_t copy_stream = nullptr; cudaEvent_t ev_final[HK_N_READOUTS] = {}, ev_copied[HK_N_READOUTS] = {};
    int async_next = 0; bool async_used[HK_N_READOUTS] = {};
    // fork / join of the per-material shading kernels of one bounce (independent queues) over side streams
    cudaStream_t shade_streams[3] = {nullptr, nullptr, nullptr}; cudaEvent_t ev_fork = nullptr, ev_join[12] = {};
    bool concurrent_shade = true; int shade_fork_slot = 0;
    bool sort_rays = false;                  // HK_SORT_RAYS=1: regroup the continuation-ray queue by direction octant before each trace (k_sort_rays); measured: no gain
    // the shadow pass of bounce b on its own stream, overlapping trace + route of bounce b+1 (opaque-only scenes)
    cudaStream_t shadow_stream = nullptr; cudaEvent_t ev_shaded = nullptr, ev_shadowed = nullptr; int bounce_par = 0;
    DevBuf b_sobol_top, b_sobol_dims, b_sobol_dimhash;        // ZSobol prefix cache (SobolParams::top)
    int32_t sobol_cache_key[6] = {0, 0, 0, 0, 0, -1};          // width, height, log2_spp, nb4, seed, cached depths
    bool sobol_cache_enabled = true;
    size_t n_slots = 0;
    HkStats stats;
    uint64_t launches = 0;
    // optional per-stage profiling (hk_set_profiling): CUDA events around every stage launch on the launching stream
    int profiling = 0;
    struct StageEv { int stage; cudaEvent_t a, b; };
    std::vector<StageEv> stage_events; size_t stage_ev_used = 0;
    double stage_ms[HK_N_STAGES]; uint64_t stage_launches[HK_N_STAGES];
    DevBuf b_work_ctr;
    // profiling bit 2: per-bounce queue counts and stage times of the most recent sample pass (host sync per bounce)
    std::vector<std::array<uint32_t, HK_N_QUEUE_COUNTERS>> bounce_counts;
    std::vector<std::array<double, HK_N_STAGES>> bounce_ms;
    // ---- frame pipelining: a second render lane ----------------------------------------------------------------------------------
    // One-sample hk_render_samples calls (the interactive render! loop: one sample + one read-out per frame) rotate over HK_N_LANES
    // lanes, each with its own path-state pool, counters, render stream, side streams and events, so that frame k+1's early bounces
    // fill the GPU while frame k's deep bounces -- a few rays per stage, each stage a latency floor -- drain.  The film is shared; its
    // accumulation order (sample order) and every read / clear of it are kept by events: ev_lane_film[l] = lane l's last
    // k_film_accumulate, ev_film_touch = the last film read-out / clear on the main stream.  Only when the library owns the main stream.
    struct AltLane {
        PathState S; DevBuf b_state, b_counts; size_t n_slots = 0;
        cudaStream_t stream = nullptr, shade_streams[3] = {nullptr, nullptr, nullptr}, shadow_stream = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join[12] = {}, ev_shaded = nullptr, ev_shadowed = nullptr, ev0 = nullptr, ev1 = nullptr;
    };
#ifndef HK_N_LANES
#define HK_N_LANES 4                   // the main lane + three more: up to four one-sample frames in flight (at most 8)
#endif
    AltLane alts[HK_N_LANES - 1];
    bool lanes_ready = false;
    cudaEvent_t ev_lane_film[HK_N_LANES] = {}, ev_film_touch = nullptr;
    bool lane_pending[HK_N_LANES] = {}, film_touch_pending = false, frame_pipeline = true;
    int next_lane = 0, last_lane = 0, last_accum_lane = -1, lane_max_count = 1;
    // swap the main lane's per-pass resources with lane k's (hk_render_samples* runs unchanged on whichever is current)
    void swap_lane(int k) {
        AltLane& alt = alts[k - 1];
        std::swap(S, alt.S); std::swap(b_state, alt.b_state); std::swap(b_counts, alt.b_counts); std::swap(n_slots, alt.n_slots);
        std::swap(stream, alt.stream); for (int i = 0; i < 3; i++) std::swap(shade_streams[i], alt.shade_streams[i]); std::swap(shadow_stream, alt.shadow_stream);
        std::swap(ev_fork, alt.ev_fork); for (int i = 0; i < 12; i++) std::swap(ev_join[i], alt.ev_join[i]);
        std::swap(ev_shaded, alt.ev_shaded); std::swap(ev_shadowed, alt.ev_shadowed); std::swap(ev0, alt.ev0); std::swap(ev1, alt.ev1);
    }
    bool any_alt_pending() const { for (int l = 1; l < HK_N_LANES; l++) if (lane_pending[l]) return true; return false; }
    void sync_alt_lanes() { for (int l = 1; l < HK_N_LANES; l++) if (lane_pending[l] && alts[l - 1].stream) { cudaStreamSynchronize(alts[l - 1].stream); lane_pending[l] = false; } }
    HkContext() { std::memset(&D, 0, sizeof(D)); std::memset(&S, 0, sizeof(S)); for (auto& a : alts) std::memset(&a.S, 0, sizeof(a.S)); std::memset(&params, 0, sizeof(params)); std::memset(&stats, 0, sizeof(stats)); }
};
// entry-point prologue: select the device and wait (on the host) for the second render lane, so that everything but the render /
// asynchronous read-out entry points sees the single-stream behaviour
static inline void hk_enter(HkContext* ctx) {
    cudaSetDevice(ctx->device);
    ctx->sync_alt_lanes();
    if (ctx->film_touch_pending) cudaStreamWaitEvent(ctx->stream, ctx->ev_film_touch, 0);      // an asynchronous read-out's finalize runs on the copy stream
}
// the stream-ordered form for the asynchronous film readers / hk_clear on the main stream: the main stream waits for the second
// lane's last film accumulation (which itself waited for everything before it)
static inline void hk_enter_film_async(HkContext* ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->last_accum_lane > 0 && ctx->lane_pending[ctx->last_accum_lane]) cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_film[ctx->last_accum_lane], 0);      // (each accumulation waited for the one before it)
    if (ctx->film_touch_pending) cudaStreamWaitEvent(ctx->stream, ctx->ev_film_touch, 0);
}
static inline void hk_film_touched(HkContext* ctx) {      // after a film read-out / clear was enqueued on the main stream
    if (ctx->ev_film_touch) { cudaEventRecord(ctx->ev_film_touch, ctx->stream); ctx->film_touch_pending = true; }
}

// build_nanovdb_from_dense on the device (hk_nvdb_build.cu): offsets / counts of the tree it wrote
struct HkNvdbBuilt { uint64_t bytes, root_off, upper_off, lower_off, leaf_off; int32_t n_up, n_low, n_leaf; int32_t idx_min[3], idx_max[3]; };
int32_t hk_nvdb_build_dense(HkContext* ctx, const float* dens_dev, const int32_t res[3], float background, DevBuf& tree, HkNvdbBuilt& out);

// (re)build the uplift cache after an upload that changes a constant colour (hk_api.cu)
extern "C" int32_t hk_refresh_uplift_cache(HkContext* ctx);
