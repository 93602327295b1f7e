/*
 * hikari_cuda_testing.h — batch entry points that run ONE device function of the VolPath path over an array of
 * inputs, so the parity tests can compare every stage with the CPU oracle in isolation (tests/test_parity_*.py).
 * They are part of libhikari_cuda.so but not of the rendering API; each mirrors an ok_test_* oracle function.
 * Also: the host-side scene-build helpers (CPU code that stays on the host in the reference as well).
 */
#ifndef HIKARI_CUDA_TESTING_H
#define HIKARI_CUDA_TESTING_H
#include "hikari_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

/* fn: 0 expf 1 logf 2 sinf 3 cosf 4 coshf 5 atanhf 6 powf(x, y) 7 log1pf of csrc/hk_detmath.h, evaluated on the DEVICE (y may be NULL
 * unless fn == 6): must equal the host evaluation of the same header bit for bit */
int32_t hk_test_detmath(HkContext* ctx, int32_t fn, const float* x, const float* y, uint64_t n, float* out);
/* q[n][4] = (px, py, sample_idx, dim) -> zsobol_sample_1d / _2d (src/sampler/sobol.jl:269-309) */
int32_t hk_test_sobol(HkContext* ctx, const int32_t* q, uint64_t n, int32_t log2_spp, int32_t n_base4_digits, uint32_t seed, float* out1d, float* out2d);
/* ray [n_slots][8] and hit [n_slots][4] (t, prim1, b1, b2) state after the last pass: per slot, the closest hit of the
 * last ray traced for that path */
int32_t hk_test_read_rays(HkContext* ctx, float* rays, float* hits, uint64_t n_slots);
/* select the Sobol' evaluation: 0 = generic matrix loop, 1 = closed forms for dimensions 0/1 (only valid when
 * hk_upload_tables verified the table structure); returns the previous mode */
int32_t hk_test_sobol_mode(HkContext* ctx, int32_t fast);
/* 0 / 1: disable / enable the upload-time uplift cache (rgb_to_spectrum of constant colours); returns the previous
 * setting.  Images must be bit-identical either way. */
int32_t hk_test_uplift_cache(HkContext* ctx, int32_t on);
/* 0 / 1: disable / enable the per-pixel ZSobol prefix cache (applies at the next hk_set_params); returns the previous
 * setting.  Images must be bit-identical either way (tests/test_parity_gpu.py). */
int32_t hk_test_sobol_cache(HkContext* ctx, int32_t on);
/* in[n][10] = p, wo, type_idx1, vec_idx1, type_idx2, vec_idx2 -> mix_hash_float (src/materials/mix-material.jl:114-158) */
int32_t hk_test_mix_hash(HkContext* ctx, const float* in, uint64_t n, float* out);
/* v[n][3] -> pbrt_hash(Vec3f), mix_bits(hash), two pcg32 floats seeded (hash, mix) (spectral-eval.jl:575-815) */
int32_t hk_test_hashes(HkContext* ctx, const float* v3, uint64_t n, uint64_t* out_hash, uint64_t* out_mix, float* out_pcg);
/* u[n] -> lambda[n][4], pdf[n][4] (src/spectral/spectral.jl:221-249) */
int32_t hk_test_wavelengths(HkContext* ctx, const float* u, uint64_t n, float* lambda, float* pdf);
/* kind 0 uplift_rgb, 1 unbounded, 2 illuminant (uplift.jl:255-308,514-538); poly = rgb_to_spectrum coefficients */
int32_t hk_test_uplift(HkContext* ctx, int32_t kind, const float* rgb, const float* lambda, uint64_t n, float* out, float* poly);
/* spectral_to_xyz + xyz_to_linear_srgb (color.jl:426-440,572-579) */
int32_t hk_test_spectral_to_rgb(HkContext* ctx, const float* L, const float* lambda, const float* pdf, uint64_t n, float* xyz, float* rgb);
/* u[n][2] -> (px, py, weight) with the filter set by hk_set_filter (filter.jl:834-953) */
int32_t hk_test_filter(HkContext* ctx, const float* u, uint64_t n, float* out);
/* camera rays of one sample pass: out[n_pixels][8] = o, d, lambda0, filter weight (volpath.jl:125-205) */
int32_t hk_test_camera_rays(HkContext* ctx, int32_t sample_idx, float* out);
/* in[n][17] = wo, n, lambda[4], u.xy, uc, regularize, wi ; out[n][16] = sample(wi,f[4],pdf,specular,eta_scale), eval(f[4],pdf), 0 */
int32_t hk_test_bsdf(HkContext* ctx, uint32_t material_idx, const float* in, uint64_t n, float* out);
/* in[n][10] = p, n, lambda_u, uc, u.xy ; out[n][16] = light idx, pmf, Li[4], wi, pdf, p_light, is_delta, pmf replay, 0 */
int32_t hk_test_lights(HkContext* ctx, const float* in, uint64_t n, float* out);
/* in[n][4] = d, lambda_u -> out[n][5] = Le[4], env pdf (lights.jl:408-467) */
int32_t hk_test_escaped(HkContext* ctx, const float* in, uint64_t n, float* out);
/* in[n][8] = o, d, t_max, lambda_u -> out[n][16] = event, beta[4], r_u[4], r_l[4], p (delta-tracking.jl:142-453) */
int32_t hk_test_delta_tracking(HkContext* ctx, uint32_t medium, const float* in, uint64_t n, float* out);
/* p[n][3] -> density of a Grid / NanoVDB medium (media.jl:1544-1595, nanovdb.jl:426-469) */
int32_t hk_test_density(HkContext* ctx, uint32_t medium, const float* p, uint64_t n, float* out);
/* in[n][8] = o, d, t_max, lambda_u -> out[n][12] = T[4], r_u[4], r_l[4] (intersection.jl:446-542) */
int32_t hk_test_ratio_tracking(HkContext* ctx, uint32_t medium, const float* in, uint64_t n, float* out);
/* node visits / triangle tests of a closest-hit batch (roofline accounting) */
int32_t hk_test_trace_counts(HkContext* ctx, const float* rays, uint64_t n, uint64_t* out_nodes, uint64_t* out_tris);
/* per-slot spectral radiance L, wavelengths, pdfs, filter weights of the LAST sample pass (n_pixels * last batch) */
int32_t hk_test_read_pass(HkContext* ctx, float* L, float* lambda, float* pdf, float* fweight, uint64_t n_slots);

/* host-side helpers (CPU): sRGB->spectrum table generator (src/spectral/rgb2spec_gen.jl:332-409) and BVH light
 * sampler construction (src/lights/bvh-light-sampler.jl:283-466) */
int32_t hk_host_generate_rgb2spec(int32_t res, const double* cie_x, const double* cie_y, const double* cie_z,
                                  const double* d65_normalised, float* out_scale, float* out_coeffs);
int32_t hk_host_build_light_sampler(const HkLight* lights, uint32_t n_lights, HkLightBVHNode* out_nodes, uint32_t* out_n_nodes,
                                    uint32_t* out_trails, int32_t* out_infinite, uint32_t* out_n_infinite, uint32_t* out_n_bvh);
/* the BVH8 builder on its own (csrc/hk_bvh.cpp; replaces Raycore's accel build): nodes are 80-byte HkBvhNode records, triangles
 * 48-byte HkBvhTri records (csrc/hk_bvh.h).  Always returns the sizes; copies when the capacities (in records) suffice. */
int32_t hk_host_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, void* out_nodes, uint64_t nodes_cap,
                           void* out_tris, uint64_t tris_cap, uint64_t* n_nodes, uint64_t* n_out_tris);
#ifdef __cplusplus
}
#endif
#endif
