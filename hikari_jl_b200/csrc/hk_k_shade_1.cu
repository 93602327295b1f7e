// hk_k_shade_1.cu — translation unit 1 of 3 of the per-material shading kernels (hk_wavefront.cuh, HK_TU_SHADE): k_shade<TYPE> for
// HK_MAT_MATTE, HK_MAT_MIRROR, HK_MAT_GLASS, HK_MAT_CONDUCTOR, HK_SHADE_MATTE_TEX, HK_MAT_THIN_DIELECTRIC, HK_MAT_DIFFUSE_TRANSMISSION.
#define HK_TU_SHADE
#include "hk_launch.h"

bool hkl_shade_1(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    switch (type) {
        case HK_MAT_MATTE: launch_shade_class<HK_MAT_MATTE>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_MIRROR: launch_shade_class<HK_MAT_MIRROR>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_GLASS: launch_shade_class<HK_MAT_GLASS>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_CONDUCTOR: launch_shade_class<HK_MAT_CONDUCTOR>(grid, st, D, S, A, next, par); return true;
        case HK_SHADE_MATTE_TEX: launch_shade_class<HK_SHADE_MATTE_TEX>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_THIN_DIELECTRIC: launch_shade_class<HK_MAT_THIN_DIELECTRIC>(grid, st, D, S, A, next, par); return true;
        case HK_MAT_DIFFUSE_TRANSMISSION: launch_shade_class<HK_MAT_DIFFUSE_TRANSMISSION>(grid, st, D, S, A, next, par); return true;
        default: return false;
    }
}
