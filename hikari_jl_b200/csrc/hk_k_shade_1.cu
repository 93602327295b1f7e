// hk_k_shade_1.cu — translation unit 1 of 3 of the per-material shading kernels (hk_wavefront.cuh, HK_TU_SHADE): k_shade<TYPE> for
// HK_MAT_MATTE, HK_MAT_MIRROR, HK_MAT_GLASS, HK_MAT_CONDUCTOR, HK_SHADE_MATTE_TEX, HK_MAT_THIN_DIELECTRIC, HK_MAT_DIFFUSE_TRANSMISSION.
#define HK_TU_SHADE
#include "hk_launch.h"

bool hkl_shade_1(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    switch (type) {
        case HK_MAT_MATTE: if (D.split_lights) k_shade<HK_MAT_MATTE, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_MATTE, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_MAT_MIRROR: if (D.split_lights) k_shade<HK_MAT_MIRROR, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_MIRROR, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_MAT_GLASS: if (D.split_lights) k_shade<HK_MAT_GLASS, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_GLASS, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_MAT_CONDUCTOR: if (D.split_lights) k_shade<HK_MAT_CONDUCTOR, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_CONDUCTOR, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_SHADE_MATTE_TEX: if (D.split_lights) k_shade<HK_SHADE_MATTE_TEX, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_SHADE_MATTE_TEX, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_MAT_THIN_DIELECTRIC: if (D.split_lights) k_shade<HK_MAT_THIN_DIELECTRIC, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_THIN_DIELECTRIC, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        case HK_MAT_DIFFUSE_TRANSMISSION: if (D.split_lights) k_shade<HK_MAT_DIFFUSE_TRANSMISSION, true><<<grid, 128, 0, st>>>(D, S, A, next, par); else k_shade<HK_MAT_DIFFUSE_TRANSMISSION, false><<<grid, 128, 0, st>>>(D, S, A, next, par); return true;
        default: return false;
    }
}
