// hk_launch.h — host-side launchers of the kernels that live in their own translation units (hk_k_trace.cu, hk_k_media.cu,
// hk_k_shade_*.cu): libhikari_cuda.so is compiled as several units in parallel (__graft_entry__.build_cuda), hk_api.cu only
// sees these prototypes.  Every launcher enqueues exactly one kernel on `st`.
#pragma once
#include "hk_wavefront.cuh"

// hk_k_trace.cu (BVH8 traversal kernels)
void hkl_trace(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int cur, int round, unsigned long long* work);
void hkl_shadow_opaque(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, unsigned long long* work, int par);
void hkl_shadow_seg_trace(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int round, unsigned long long* work);
void hkl_trace_batch(bool any, bool count, int grid, cudaStream_t st, const DevBvh& B, const float4* rays, uint32_t n, float4* hits, uint8_t* occluded,
                     uint32_t* cursor, unsigned long long* counters);
void hkl_aux_buffers(int grid, cudaStream_t st, const DevScene& D, float* albedo, float* normal, float* depth, float miss_depth);
void hkl_detect_camera_medium(cudaStream_t st, const DevScene& D, uint32_t* out);
// hk_k_media.cu (delta / ratio tracking)
void hkl_medium_track(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S);
void hkl_medium_finish(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next);
void hkl_shadow_seg_ratio(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int round);
// hk_k_lights.cu (emissive-hit MIS + NEE light sample of every surface hit)
void hkl_hit_lights(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A);
void hkl_hit_lights_bvh(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A);
// hk_k_shade_{1,2,3}.cu: each handles a subset of the shading classes and returns false for the others
bool hkl_shade_1(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par);
bool hkl_shade_2(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par);
bool hkl_shade_3(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par);
inline void hkl_shade(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    if (!hkl_shade_1(type, grid, st, D, S, A, next, par) && !hkl_shade_2(type, grid, st, D, S, A, next, par)) hkl_shade_3(type, grid, st, D, S, A, next, par);
}
