// hk_k_media.cu — translation unit of the participating-media kernels (hk_wavefront.cuh, HK_TU_MEDIA): k_medium_track,
// k_medium_finish, k_shadow_seg_ratio.
#define HK_TU_MEDIA
#include "hk_launch.h"

void hkl_medium_track(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S) {
    if (rgb) k_medium_track<true><<<grid, 128, 0, st>>>(D, S);
    else k_medium_track<false><<<grid, 128, 0, st>>>(D, S);
}
void hkl_medium_finish(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next) {
    k_medium_finish<<<grid, 128, 0, st>>>(D, S, A, next);
}
void hkl_shadow_seg_ratio(bool rgb, int grid, cudaStream_t st, const DevScene& D, const PathState& S, int round) {
    if (rgb) k_shadow_seg_ratio<true><<<grid, 128, 0, st>>>(D, S, round);
    else k_shadow_seg_ratio<false><<<grid, 128, 0, st>>>(D, S, round);
}
