// hk_k_trace.cu — translation unit of the BVH8 traversal kernels (hk_wavefront.cuh, HK_TU_TRACE): k_trace, k_shadow_opaque,
// k_shadow_seg_trace, k_trace_batch, k_aux_buffers, k_detect_camera_medium.
#define HK_TU_TRACE
#include "hk_launch.h"

// kernel<A, B><<<cfg>>>(args) with two run-time booleans
#define HK_DISPATCH2(a, b, kernel, cfg, args) do { \
    if (a) { if (b) kernel<true, true><<<HK_UNPACK cfg>>>args; else kernel<true, false><<<HK_UNPACK cfg>>>args; } \
    else { if (b) kernel<false, true><<<HK_UNPACK cfg>>>args; else kernel<false, false><<<HK_UNPACK cfg>>>args; } } while (0)
#define HK_UNPACK(...) __VA_ARGS__

void hkl_trace(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int cur, int round, unsigned long long* work) {
    HK_DISPATCH2(count, D.bvh.inst != nullptr, k_trace, (grid, HK_TRACE_THREADS, 0, st), (D, S, cur, round, work));
}
void hkl_shadow_opaque(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, unsigned long long* work, int par) {
    HK_DISPATCH2(count, D.bvh.inst != nullptr, k_shadow_opaque, (grid, HK_TRACE_THREADS, 0, st), (D, S, work, par));
}
void hkl_shadow_seg_trace(bool count, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int round, unsigned long long* work) {
    HK_DISPATCH2(count, D.bvh.inst != nullptr, k_shadow_seg_trace, (grid, HK_TRACE_THREADS, 0, st), (D, S, round, work));
}
void hkl_trace_batch(bool any, bool count, int grid, cudaStream_t st, const DevBvh& B, const float4* rays, uint32_t n, float4* hits, uint8_t* occluded,
                     uint32_t* cursor, unsigned long long* counters) {
    const bool inst = B.inst != nullptr;
    if (any) { if (inst) k_trace_batch<true, false, true><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); else k_trace_batch<true, false, false><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); }
    else if (count) { if (inst) k_trace_batch<false, true, true><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); else k_trace_batch<false, true, false><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); }
    else { if (inst) k_trace_batch<false, false, true><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); else k_trace_batch<false, false, false><<<grid, HK_TRACE_THREADS, 0, st>>>(B, rays, n, hits, occluded, cursor, counters); }
}
void hkl_aux_buffers(int grid, cudaStream_t st, const DevScene& D, float* albedo, float* normal, float* depth, float miss_depth) {
    if (D.bvh.inst) k_aux_buffers<true><<<grid, HK_TRACE_THREADS, 0, st>>>(D, albedo, normal, depth, miss_depth);
    else k_aux_buffers<false><<<grid, HK_TRACE_THREADS, 0, st>>>(D, albedo, normal, depth, miss_depth);
}
void hkl_detect_camera_medium(cudaStream_t st, const DevScene& D, uint32_t* out) {
    if (D.bvh.inst) k_detect_camera_medium<true><<<1, 32, 0, st>>>(D, out); else k_detect_camera_medium<false><<<1, 32, 0, st>>>(D, out);
}
