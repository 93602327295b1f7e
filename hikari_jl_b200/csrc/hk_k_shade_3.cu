// hk_k_shade_3.cu — translation unit 3 of 3 of the per-material shading kernels (hk_wavefront.cuh, HK_TU_SHADE): k_shade<TYPE> for
// HK_MAT_COATED_DIFFUSE_TRANSMISSION.
#define HK_TU_SHADE
// the coated / layered kernels are several times the 32 KB instruction cache: the light-sampling helpers are real functions here
// (measured on C5: shading 15.9 -> 12.8 ms; in the small kernels of hk_k_shade_1.cu the same switch costs 13-16 %)
#define HK_NOINLINE_LIGHTS 1
// ... and so are Spec / float and norm3 (C5 shading 11.84 -> 11.48 ms/step; ncu: the coated-diffuse kernel is 196 KB of SASS and
// spends 18 warps per issue in no_instruction, profiles/r02_c5_k_shade_coated_ncu_summary.txt)
#ifndef HK_LAYERED_SPDIV
#define HK_LAYERED_SPDIV 1
#endif
#ifndef HK_LAYERED_VEC
#define HK_LAYERED_VEC 1
#endif
#ifndef HK_LAYERED_SOBOL
#define HK_LAYERED_SOBOL 0
#endif
#define HK_NOINLINE_SPDIV HK_LAYERED_SPDIV
#define HK_NOINLINE_VEC HK_LAYERED_VEC
#define HK_NOINLINE_SOBOL HK_LAYERED_SOBOL
#include "hk_launch.h"

bool hkl_shade_3(int type, int grid, cudaStream_t st, const DevScene& D, const PathState& S, const PassArgs& A, int next, int par) {
    switch (type) {
        case HK_MAT_COATED_DIFFUSE_TRANSMISSION: launch_shade_class<HK_MAT_COATED_DIFFUSE_TRANSMISSION>(grid, st, D, S, A, next, par); return true;
        default: return false;
    }
}
