// hk_k_shade_2.cu — translation unit 2 of 3 of the per-material shading kernels (hk_wavefront.cuh, HK_TU_SHADE): k_shade<TYPE> for
// HK_MAT_COATED_DIFFUSE, HK_MAT_COATED_CONDUCTOR.
#define HK_TU_SHADE
#include "hk_launch.h"

bool hkl_shade_2(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    switch (type) {
        case HK_MAT_COATED_DIFFUSE: launch_shade_class<HK_MAT_COATED_DIFFUSE>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_COATED_CONDUCTOR: launch_shade_class<HK_MAT_COATED_CONDUCTOR>(grid, st, D, S, A, next, par); return true;
        default: return false;
    }
}
