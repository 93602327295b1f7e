// hk_k_media.cu — translation unit of the participating-media kernels (hk_wavefront.cuh, HK_TU_MEDIA): k_medium_track,
// k_medium_finish, k_shadow_seg_ratio.
#define HK_TU_MEDIA
// The tracking kernels are ~70 KB of SASS against a 32 KB instruction cache (ncu: no_instruction is their second largest stall).
// Spec / float -- four IEEE divisions, ~40 instructions at each of ~14 call sites -- is a real function in this translation unit
// (C4 +4 %; elsewhere inlining it is the faster choice, hk_math.cuh).
#ifndef HK_NOINLINE_SPDIV
#define HK_NOINLINE_SPDIV 1
#endif
#include "hk_launch.h"

void hkl_medium_track(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S) {
    const size_t sm = 4 * ((size_t)D.smem_mask_words + HK_LC_WORDS * 128);      // the staged empty-cell mask (<= 32 KB) + one leaf cache per lane
    if (rgb) k_medium_track<true><<<grid, 128, sm, st>>>(D, S);
    else k_medium_track<false><<<grid, 128, sm, st>>>(D, S);
}
void hkl_medium_finish(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next) {
    k_medium_finish<<<grid, 128, 0, st>>>(D, S, A, next);
}
void hkl_shadow_seg_ratio(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int round) {
    const size_t sm = 4 * ((size_t)D.smem_mask_words + HK_LC_WORDS * 128);
    if (rgb) k_shadow_seg_ratio<true><<<grid, 128, sm, st>>>(D, S, round);
    else k_shadow_seg_ratio<false><<<grid, 128, sm, st>>>(D, S, round);
}

#ifdef HK_MEDIA_STATS
// development: print and reset the tracking statistics (hk_media.cuh)
extern "C" void hk_dev_media_stats(unsigned long long* out16) { cudaDeviceSynchronize(); cudaMemcpyFromSymbol(out16, g_media_stats, sizeof(unsigned long long) * 32); unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_media_stats, z, sizeof(z)); }
#endif
