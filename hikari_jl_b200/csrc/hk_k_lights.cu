// hk_k_lights.cu — translation unit of k_hit_lights (hk_wavefront.cuh, HK_TU_LIGHTS): emissive-hit MIS and the NEE light sample
// (light-BVH selection + sample_light) of every surface hit of a bounce, for all material queues at once.
#define HK_TU_LIGHTS
#include "hk_launch.h"

void hkl_hit_lights(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A) {
    k_hit_lights<<<grid, 128, 0, st>>>(D, S, A);
}
void hkl_hit_lights_bvh(int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A) {
#if HK_LIGHTS_COMPACT
    k_hit_lights_bvh<<<grid, 128, 0, st>>>(D, S, A);
#endif
}
