// hk_k_shade_3.cu — translation unit 3 of 3 of the per-material shading kernels (hk_wavefront.cuh, HK_TU_SHADE): k_shade<TYPE> for
// HK_MAT_COATED_DIFFUSE_TRANSMISSION.
#define HK_TU_SHADE
#include "hk_launch.h"

bool hkl_shade_3(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    switch (type) {
        case HK_MAT_COATED_DIFFUSE_TRANSMISSION: launch_shade_class<HK_MAT_COATED_DIFFUSE_TRANSMISSION>(grid, st, D, S, A, next, par); return true;
        default: return false;
    }
}
