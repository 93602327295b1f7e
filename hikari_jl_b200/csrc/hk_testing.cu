// hk_testing.cu — batch kernels behind include/hikari_cuda_testing.h: each runs ONE device function of the path
// over an input array so tests can compare it with the CPU oracle in isolation.  Its own translation unit of libhikari_cuda.so.
#include "hk_context.h"
#include <cstring>

#define TK_LOOP(n) for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (n); i += (uint64_t)gridDim.x * blockDim.x)

__global__ void tk_detmath(int fn, const float* __restrict__ x, const float* __restrict__ y, uint64_t n, float* out) {
    TK_LOOP(n) {
        const float a = x[i], b = y ? y[i] : 0.0f;
        out[i] = fn == 0 ? dm_expf(a) : fn == 1 ? dm_logf(a) : fn == 2 ? dm_sinf(a) : fn == 3 ? dm_cosf(a) : fn == 4 ? dm_coshf(a) : fn == 5 ? dm_atanhf(a) : fn == 6 ? dm_powf(a, b) : dm_log1pf(a);
    }
}
__global__ void tk_sobol(SobolParams P, const int32_t* __restrict__ q, uint64_t n, float* o1, float* o2) {
    TK_LOOP(n) {
        const int32_t* e = q + 4 * i;
        o1[i] = zsobol_1d(P, e[0], e[1], e[2], e[3]);
        float2 v = zsobol_2d(P, e[0], e[1], e[2], e[3]);
        o2[2 * i] = v.x; o2[2 * i + 1] = v.y;
    }
}
__global__ void tk_mix_hash(const float* __restrict__ in, uint64_t n, float* out) {
    TK_LOOP(n) {
        const float* e = in + 10 * i;
        out[i] = mix_hash_float(f3(e[0], e[1], e[2]), f3(e[3], e[4], e[5]), (uint32_t)e[6], (uint32_t)e[7], (uint32_t)e[8], (uint32_t)e[9]);
    }
}
__global__ void tk_hashes(const float* __restrict__ v, uint64_t n, uint64_t* oh, uint64_t* om, float* op) {
    TK_LOOP(n) {
        uint64_t h = hash_f3(f3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
        uint64_t m = mix_bits(h);
        oh[i] = h; om[i] = m;
        Pcg32 r = pcg32_init(h, m);
        op[2 * i] = pcg32_f32(r); op[2 * i + 1] = pcg32_f32(r);
    }
}
__global__ void tk_wavelengths(const float* __restrict__ u, uint64_t n, float4* lam, float4* pdf) {
    TK_LOOP(n) { float4 l, p; sample_wavelengths_visible(u[i], l, p); lam[i] = l; pdf[i] = p; }
}
__global__ void tk_uplift(DevTables T, int kind, const float* __restrict__ rgb, const float4* __restrict__ lam, uint64_t n, float4* out, float* poly) {
    TK_LOOP(n) {
        float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        out[i] = kind == 0 ? uplift_rgb(T, r, g, b, lam[i]) : (kind == 1 ? uplift_rgb_unbounded(T, r, g, b, lam[i]) : uplift_rgb_illuminant(T, r, g, b, lam[i]));
        Poly3 p = rgb_to_spectrum(T, r, g, b);
        poly[3 * i] = p.c0; poly[3 * i + 1] = p.c1; poly[3 * i + 2] = p.c2;
    }
}
__global__ void tk_spec2rgb(DevTables T, const float4* __restrict__ L, const float4* __restrict__ lam, const float4* __restrict__ pdf, uint64_t n, float* xyz, float* rgb) {
    TK_LOOP(n) {
        float3 x = spectral_to_xyz(T, L[i], lam[i], pdf[i]);
        float3 r = xyz_to_linear_srgb(x);
        xyz[3 * i] = x.x; xyz[3 * i + 1] = x.y; xyz[3 * i + 2] = x.z; rgb[3 * i] = r.x; rgb[3 * i + 1] = r.y; rgb[3 * i + 2] = r.z;
    }
}
__global__ void tk_filter(DevFilter F, const float* __restrict__ u, uint64_t n, float* out) {
    TK_LOOP(n) { float px, py, w; filter_sample(F, make_float2(u[2 * i], u[2 * i + 1]), px, py, w); out[3 * i] = px; out[3 * i + 1] = py; out[3 * i + 2] = w; }
}
__global__ void tk_camera(const __grid_constant__ DevScene D, int sample_idx, float* out) {
    const uint64_t n = (uint64_t)D.width * D.height;
    TK_LOOP(n) {
        int x = (int)(i % (uint64_t)D.width) + 1, y = (int)(i / (uint64_t)D.width) + 1;
        float wu = zsobol_1d(D.sobol, x, y, sample_idx, 1);
        float2 j = zsobol_2d(D.sobol, x, y, sample_idx, 3);
        float2 lens = zsobol_2d(D.sobol, x, y, sample_idx, 6);
        float fx, fy, fw; filter_sample(D.filter, j, fx, fy, fw);
        float4 lam, pdf; sample_wavelengths_visible(wu, lam, pdf);
        float3 o, d; camera_generate_ray(D.camera, (float)x + 0.5f + fx, (float)D.height - (float)y + 1.0f + 0.5f + fy, lens, o, d);
        float* r = out + 8 * i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z; r[6] = lam.x; r[7] = fw;
    }
}
template <int TYPE>
__device__ void tk_bsdf_one(const MatCtx& MC, const HkMaterial& m, const float* e, float* o) {
    float3 wo = f3(e[0], e[1], e[2]), nn = f3(e[3], e[4], e[5]);
    float4 lam = make_float4(e[6], e[7], e[8], e[9]);
    BsdfSample bs = sample_bsdf<TYPE>(MC, m, wo, nn, lam, make_float2(e[10], e[11]), e[12], e[13] != 0.0f);
    o[0] = bs.wi.x; o[1] = bs.wi.y; o[2] = bs.wi.z; o[3] = bs.f.x; o[4] = bs.f.y; o[5] = bs.f.z; o[6] = bs.f.w;
    o[7] = bs.pdf; o[8] = bs.specular ? 1.0f : 0.0f; o[9] = bs.eta_scale;
    BsdfEval be = eval_bsdf<TYPE>(MC, m, wo, f3(e[14], e[15], e[16]), nn, lam);
    o[10] = be.f.x; o[11] = be.f.y; o[12] = be.f.z; o[13] = be.f.w; o[14] = be.pdf; o[15] = 0.0f;
}
__global__ void __launch_bounds__(128) tk_bsdf(const __grid_constant__ DevScene D, uint32_t mat_idx, const float* __restrict__ in, uint64_t n, float* out) {
    MatCtx MC = mat_ctx(D);
    const HkMaterial& m = D.materials[mat_idx - 1];
    TK_LOOP(n) {
        const float* e = in + 17 * i; float* o = out + 16 * i;
        switch (m.type) {
            case HK_MAT_MATTE: tk_bsdf_one<HK_MAT_MATTE>(MC, m, e, o); break;
            case HK_MAT_MIRROR: tk_bsdf_one<HK_MAT_MIRROR>(MC, m, e, o); break;
            case HK_MAT_GLASS: tk_bsdf_one<HK_MAT_GLASS>(MC, m, e, o); break;
            case HK_MAT_CONDUCTOR: tk_bsdf_one<HK_MAT_CONDUCTOR>(MC, m, e, o); break;
            case HK_MAT_COATED_DIFFUSE: tk_bsdf_one<HK_MAT_COATED_DIFFUSE>(MC, m, e, o); break;
            case HK_MAT_THIN_DIELECTRIC: tk_bsdf_one<HK_MAT_THIN_DIELECTRIC>(MC, m, e, o); break;
            case HK_MAT_DIFFUSE_TRANSMISSION: tk_bsdf_one<HK_MAT_DIFFUSE_TRANSMISSION>(MC, m, e, o); break;
            case HK_MAT_COATED_CONDUCTOR: tk_bsdf_one<HK_MAT_COATED_CONDUCTOR>(MC, m, e, o); break;
            case HK_MAT_COATED_DIFFUSE_TRANSMISSION: tk_bsdf_one<HK_MAT_COATED_DIFFUSE_TRANSMISSION>(MC, m, e, o); break;
        }
    }
}
__global__ void tk_lights(const __grid_constant__ DevScene D, const float* __restrict__ in, uint64_t n, float* out) {
    LightCtx LC = light_ctx(D);
    TK_LOOP(n) {
        const float* e = in + 10 * i; float* o = out + 16 * i;
        for (int k = 0; k < 16; k++) o[k] = 0.0f;
        float3 p = f3(e[0], e[1], e[2]), nn = f3(e[3], e[4], e[5]);
        float4 lam, pdf; sample_wavelengths_visible(e[6], lam, pdf);
        float pmf; int li = bvh_sample_light_coop(LC, p, nn, e[7], pmf);      // the cooperative form the shading kernels use
        o[0] = (float)li; o[1] = pmf;
        if (li >= 1 && li <= D.n_lights) {
            LightSample ls = sample_light(LC, D.lights[li - 1], p, lam, make_float2(e[8], e[9]));
            o[2] = ls.Li.x; o[3] = ls.Li.y; o[4] = ls.Li.z; o[5] = ls.Li.w; o[6] = ls.wi.x; o[7] = ls.wi.y; o[8] = ls.wi.z; o[9] = ls.pdf;
            o[10] = ls.p_light.x; o[11] = ls.p_light.y; o[12] = ls.p_light.z; o[13] = ls.delta ? 1.0f : 0.0f;
            o[14] = bvh_light_pmf(LC, p, nn, li);
        }
    }
}
__global__ void tk_escaped(const __grid_constant__ DevScene D, const float* __restrict__ in, uint64_t n, float* out) {
    LightCtx LC = light_ctx(D);
    TK_LOOP(n) {
        const float* e = in + 4 * i;
        float3 d = f3(e[0], e[1], e[2]);
        float4 lam, pdf; sample_wavelengths_visible(e[3], lam, pdf);
        Spec Le = escaped_Le(LC, d, lam);
        out[5 * i] = Le.x; out[5 * i + 1] = Le.y; out[5 * i + 2] = Le.z; out[5 * i + 3] = Le.w; out[5 * i + 4] = env_light_pdf(LC, d);
    }
}
__global__ void __launch_bounds__(128) tk_delta(const __grid_constant__ DevScene D, uint32_t medium, const float* __restrict__ in, uint64_t n, float* out) {
    MediaCtx MC = media_ctx(D);
    TK_LOOP(n) {
        const float* e = in + 8 * i; float* o = out + 16 * i;
        float4 lam, pdf; sample_wavelengths_visible(e[7], lam, pdf);
        DeltaOut r = delta_track(MC, (int)medium, f3(e[0], e[1], e[2]), f3(e[3], e[4], e[5]), e[6], lam, sp(1.0f), sp(1.0f), sp(1.0f), 0, 1 << 30);
        o[0] = r.event == HK_EV_ABSORBED ? 0.0f : (r.event == HK_EV_SCATTER ? 1.0f : 2.0f);
        o[1] = r.beta.x; o[2] = r.beta.y; o[3] = r.beta.z; o[4] = r.beta.w; o[5] = r.r_u.x; o[6] = r.r_u.y; o[7] = r.r_u.z; o[8] = r.r_u.w;
        o[9] = r.r_l.x; o[10] = r.r_l.y; o[11] = r.r_l.z; o[12] = r.r_l.w;
        bool sc = r.event == HK_EV_SCATTER;
        o[13] = sc ? r.p.x : 0.0f; o[14] = sc ? r.p.y : 0.0f; o[15] = sc ? r.p.z : 0.0f;
    }
}
__global__ void tk_density(const __grid_constant__ DevScene D, uint32_t medium, const float* __restrict__ p, uint64_t n, float* out) {
    const DevMedium& M = D.media[medium - 1];
    TK_LOOP(n) out[i] = medium_density(M, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
__global__ void __launch_bounds__(128) tk_ratio(const __grid_constant__ DevScene D, uint32_t medium, const float* __restrict__ in, uint64_t n, float* out) {
    MediaCtx MC = media_ctx(D);
    TK_LOOP(n) {
        const float* e = in + 8 * i; float* o = out + 12 * i;
        float4 lam, pdf; sample_wavelengths_visible(e[7], lam, pdf);
        Spec T, ru, rl;
        ratio_track(MC, (int)medium, f3(e[0], e[1], e[2]), f3(e[3], e[4], e[5]), e[6], lam, T, ru, rl);
        o[0] = T.x; o[1] = T.y; o[2] = T.z; o[3] = T.w; o[4] = ru.x; o[5] = ru.y; o[6] = ru.z; o[7] = ru.w; o[8] = rl.x; o[9] = rl.y; o[10] = rl.z; o[11] = rl.w;
    }
}

// ---- host wrappers: upload inputs, launch, download outputs ------------------------------------------------------------
struct TkIO {
    HkContext* ctx; std::vector<DevBuf> bufs; int32_t rc = HK_OK;
    explicit TkIO(HkContext* c) : ctx(c) { bufs.reserve(16); }
    void* in(const void* h, size_t bytes) { bufs.emplace_back(); if (bufs.back().upload(h, bytes) != cudaSuccess) rc = HK_ERR_CUDA; return bufs.back().p; }
    void* out(size_t bytes) { bufs.emplace_back(); if (bufs.back().alloc(bytes) != cudaSuccess) rc = HK_ERR_CUDA; else cudaMemset(bufs.back().p, 0, bytes ? bytes : 16); return bufs.back().p; }
    int32_t get(void* h, const void* d, size_t bytes) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess && bytes) e = cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = HK_ERR_CUDA; }
        return rc;
    }
    ~TkIO() { for (auto& b : bufs) b.release(); }
};
#define TK_GRID(n) ((int)std::min<uint64_t>(((n) + 127) / 128 ? ((n) + 127) / 128 : 1, 148 * 16)), 128

extern "C" {
int32_t hk_test_detmath(HkContext* ctx, int32_t fn, const float* x, const float* y, uint64_t n, float* out) {
    if (!ctx || !x || !out || fn < 0 || fn > 7 || (fn == 6 && !y)) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* dx = io.in(x, 4 * n); void* dy = y ? io.in(y, 4 * n) : nullptr; float* dout = (float*)io.out(4 * n);
    if (io.rc) return io.rc;
    tk_detmath<<<TK_GRID(n)>>>(fn, (const float*)dx, (const float*)dy, n, dout); ctx->launches++;
    return io.get(out, dout, 4 * n);
}
int32_t hk_test_sobol(HkContext* ctx, const int32_t* q, uint64_t n, int32_t l2, int32_t nb4, uint32_t seed, float* o1, float* o2) {
    if (!ctx || !ctx->have_tables) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    SobolParams P{ctx->D.T.sobol, l2, nb4, seed, ctx->D.sobol.fast, nullptr, nullptr, 0u, 0};
    void* dq = io.in(q, 16 * n); float* d1 = (float*)io.out(4 * n); float* d2 = (float*)io.out(8 * n);
    if (io.rc) return io.rc;
    tk_sobol<<<TK_GRID(n)>>>(P, (const int32_t*)dq, n, d1, d2); ctx->launches++;
    io.get(o1, d1, 4 * n); return io.get(o2, d2, 8 * n);
}
// 0: generic table loop, 1: closed forms for dimensions 0/1; returns the previous mode (tests run both)
int32_t hk_test_sobol_mode(HkContext* ctx, int32_t fast) { if (!ctx) return HK_ERR_INVALID; int32_t old = ctx->D.sobol.fast; ctx->D.sobol.fast = fast ? 1 : 0; return old; }
// 0 / 1: disable / enable the uplift cache (DevTables::mat_pre ...); returns the previous setting.  Takes effect at once.
int32_t hk_test_uplift_cache(HkContext* ctx, int32_t on) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx);
    int32_t old = ctx->uplift_cache_enabled ? 1 : 0; ctx->uplift_cache_enabled = on != 0;
    int32_t rc = hk_refresh_uplift_cache(ctx);
    return rc != HK_OK ? rc : old;
}
// 0 / 1: disable / enable the ZSobol prefix cache (takes effect at the next hk_set_params); returns the previous setting
int32_t hk_test_sobol_cache(HkContext* ctx, int32_t on) { if (!ctx) return HK_ERR_INVALID; int32_t old = ctx->sobol_cache_enabled ? 1 : 0; ctx->sobol_cache_enabled = on != 0; ctx->sobol_cache_key[5] = -1; return old; }
int32_t hk_test_mix_hash(HkContext* ctx, const float* in, uint64_t n, float* out) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 40 * n); float* dout = (float*)io.out(4 * n);
    if (io.rc) return io.rc;
    tk_mix_hash<<<TK_GRID(n)>>>((const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 4 * n);
}
int32_t hk_test_hashes(HkContext* ctx, const float* v, uint64_t n, uint64_t* oh, uint64_t* om, float* op) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* dv = io.in(v, 12 * n); uint64_t* dh = (uint64_t*)io.out(8 * n); uint64_t* dm = (uint64_t*)io.out(8 * n); float* dp = (float*)io.out(8 * n);
    if (io.rc) return io.rc;
    tk_hashes<<<TK_GRID(n)>>>((const float*)dv, n, dh, dm, dp); ctx->launches++;
    io.get(oh, dh, 8 * n); io.get(om, dm, 8 * n); return io.get(op, dp, 8 * n);
}
int32_t hk_test_wavelengths(HkContext* ctx, const float* u, uint64_t n, float* lam, float* pdf) {
    if (!ctx) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* du = io.in(u, 4 * n); float4* dl = (float4*)io.out(16 * n); float4* dp = (float4*)io.out(16 * n);
    if (io.rc) return io.rc;
    tk_wavelengths<<<TK_GRID(n)>>>((const float*)du, n, dl, dp); ctx->launches++;
    io.get(lam, dl, 16 * n); return io.get(pdf, dp, 16 * n);
}
int32_t hk_test_uplift(HkContext* ctx, int32_t kind, const float* rgb, const float* lam, uint64_t n, float* out, float* poly) {
    if (!ctx || !ctx->have_tables) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* dr = io.in(rgb, 12 * n); void* dl = io.in(lam, 16 * n); float4* dout = (float4*)io.out(16 * n); float* dpoly = (float*)io.out(12 * n);
    if (io.rc) return io.rc;
    tk_uplift<<<TK_GRID(n)>>>(ctx->D.T, kind, (const float*)dr, (const float4*)dl, n, dout, dpoly); ctx->launches++;
    io.get(out, dout, 16 * n); return io.get(poly, dpoly, 12 * n);
}
int32_t hk_test_spectral_to_rgb(HkContext* ctx, const float* L, const float* lam, const float* pdf, uint64_t n, float* xyz, float* rgb) {
    if (!ctx || !ctx->have_tables) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* dL = io.in(L, 16 * n); void* dl = io.in(lam, 16 * n); void* dp = io.in(pdf, 16 * n); float* dx = (float*)io.out(12 * n); float* dr = (float*)io.out(12 * n);
    if (io.rc) return io.rc;
    tk_spec2rgb<<<TK_GRID(n)>>>(ctx->D.T, (const float4*)dL, (const float4*)dl, (const float4*)dp, n, dx, dr); ctx->launches++;
    io.get(xyz, dx, 12 * n); return io.get(rgb, dr, 12 * n);
}
int32_t hk_test_filter(HkContext* ctx, const float* u, uint64_t n, float* out) {
    if (!ctx || !ctx->have_filter) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* du = io.in(u, 8 * n); float* dout = (float*)io.out(12 * n);
    if (io.rc) return io.rc;
    tk_filter<<<TK_GRID(n)>>>(ctx->D.filter, (const float*)du, n, dout); ctx->launches++;
    return io.get(out, dout, 12 * n);
}
int32_t hk_test_camera_rays(HkContext* ctx, int32_t sample_idx, float* out) {
    if (!ctx || !ctx->have_tables || !ctx->have_cam || !ctx->have_filter || !ctx->have_params) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    uint64_t n = (uint64_t)ctx->params.width * ctx->params.height;
    float* dout = (float*)io.out(32 * n);
    if (io.rc) return io.rc;
    tk_camera<<<TK_GRID(n)>>>(ctx->D, sample_idx, dout); ctx->launches++;
    return io.get(out, dout, 32 * n);
}
int32_t hk_test_bsdf(HkContext* ctx, uint32_t mat, const float* in, uint64_t n, float* out) {
    if (!ctx || !ctx->have_tables || !ctx->have_mats) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 68 * n); float* dout = (float*)io.out(64 * n);
    if (io.rc) return io.rc;
    tk_bsdf<<<TK_GRID(n)>>>(ctx->D, mat, (const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 64 * n);
}
int32_t hk_test_lights(HkContext* ctx, const float* in, uint64_t n, float* out) {
    if (!ctx || !ctx->have_tables || !ctx->have_lights) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 40 * n); float* dout = (float*)io.out(64 * n);
    if (io.rc) return io.rc;
    tk_lights<<<TK_GRID(n)>>>(ctx->D, (const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 64 * n);
}
int32_t hk_test_escaped(HkContext* ctx, const float* in, uint64_t n, float* out) {
    if (!ctx || !ctx->have_tables || !ctx->have_lights) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 16 * n); float* dout = (float*)io.out(20 * n);
    if (io.rc) return io.rc;
    tk_escaped<<<TK_GRID(n)>>>(ctx->D, (const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 20 * n);
}
int32_t hk_test_delta_tracking(HkContext* ctx, uint32_t medium, const float* in, uint64_t n, float* out) {
    if (!ctx || !ctx->have_tables || medium < 1 || (int32_t)medium > ctx->D.n_media) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 32 * n); float* dout = (float*)io.out(64 * n);
    if (io.rc) return io.rc;
    tk_delta<<<TK_GRID(n)>>>(ctx->D, medium, (const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 64 * n);
}
int32_t hk_test_density(HkContext* ctx, uint32_t medium, const float* p, uint64_t n, float* out) {
    if (!ctx || medium < 1 || (int32_t)medium > ctx->D.n_media) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* dp = io.in(p, 12 * n); float* dout = (float*)io.out(4 * n);
    if (io.rc) return io.rc;
    tk_density<<<TK_GRID(n)>>>(ctx->D, medium, (const float*)dp, n, dout); ctx->launches++;
    return io.get(out, dout, 4 * n);
}
int32_t hk_test_ratio_tracking(HkContext* ctx, uint32_t medium, const float* in, uint64_t n, float* out) {
    if (!ctx || !ctx->have_tables || medium < 1 || (int32_t)medium > ctx->D.n_media) return HK_ERR_INVALID;
    hk_enter(ctx); TkIO io(ctx);
    void* di = io.in(in, 32 * n); float* dout = (float*)io.out(48 * n);
    if (io.rc) return io.rc;
    tk_ratio<<<TK_GRID(n)>>>(ctx->D, medium, (const float*)di, n, dout); ctx->launches++;
    return io.get(out, dout, 48 * n);
}
// rays [n_slots][8] = (o, d, t_max, 0) and hits [n_slots][4] = (t, prim1 bits, b1, b2) as the last pass left them: for
// every slot the hit record is the closest hit of the LAST ray traced for that path (full-size traversal parity checks)
int32_t hk_test_read_rays(HkContext* ctx, float* rays, float* hits, uint64_t n_slots) {
    if (!ctx || !ctx->have_params || !rays || !hits) return HK_ERR_INVALID;
    hk_enter(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    const PathState& PS = ctx->last_lane > 0 ? ctx->alts[ctx->last_lane - 1].S : ctx->S;      // the lane that rendered last (frame pipelining)
    if (n_slots > (ctx->last_lane > 0 ? ctx->alts[ctx->last_lane - 1].n_slots : ctx->n_slots)) return HK_ERR_INVALID;
    std::vector<float4> a(n_slots), b(n_slots);
    CK(cudaMemcpy(a.data(), PS.ray_a, 16 * n_slots, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), PS.ray_b, 16 * n_slots, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < n_slots; i++) {
        float* r = rays + 8 * i;
        r[0] = a[i].x; r[1] = a[i].y; r[2] = a[i].z; r[3] = a[i].w; r[4] = b[i].x; r[5] = b[i].y; r[6] = b[i].z; r[7] = 0.0f;
    }
    CK(cudaMemcpy(hits, PS.hit, 16 * n_slots, cudaMemcpyDeviceToHost));
    uint32_t* hb = reinterpret_cast<uint32_t*>(hits);
    for (uint64_t i = 0; i < n_slots; i++) hb[4 * i + 1] &= HK_PRIM_MASK;
    return HK_OK;
}
int32_t hk_test_read_pass(HkContext* ctx, float* L, float* lam, float* pdf, float* fw, uint64_t n_slots) {
    if (!ctx || !ctx->have_params) return HK_ERR_INVALID;
    hk_enter(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    const PathState& PS = ctx->last_lane > 0 ? ctx->alts[ctx->last_lane - 1].S : ctx->S;
    if (n_slots > (ctx->last_lane > 0 ? ctx->alts[ctx->last_lane - 1].n_slots : ctx->n_slots)) return HK_ERR_INVALID;
    CK(cudaMemcpy(L, PS.L, 16 * n_slots, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(lam, PS.lambda, 16 * n_slots, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pdf, PS.lpdf, 16 * n_slots, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(fw, PS.fweight, 4 * n_slots, cudaMemcpyDeviceToHost));
    return HK_OK;
}
}  // extern "C"
