// hk_denoise.cuh — edge-avoiding a-trous wavelet denoiser over the film's auxiliary buffers.
// Reference: src/denoise.jl (weights :66-114, 5x5 a-trous pass :123-207, 3x3 luminance variance :216-258, driver :301-372).
// All images are (H, W) column-major like film.framebuffer: idx -> row = idx % H, col = idx / H, so consecutive threads read
// consecutive rows of one column (coalesced) and the 25 taps of a pass come from L1 / L2.  HBM-bound: 12 B in + 12 B out per
// pixel and pass plus the normal / depth / variance planes.
#pragma once
#include "hk_math.cuh"

HK_DEV float dn_lum(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }      // :66-68

__global__ void __launch_bounds__(256) k_denoise_variance(float* __restrict__ variance, const float* __restrict__ in, int W, int H) {      // :216-258
    const uint32_t n = (uint32_t)W * (uint32_t)H;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int row = (int)(idx % (uint32_t)H), col = (int)(idx / (uint32_t)H);
        float s = 0.0f, s2 = 0.0f; int count = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int qr = row + dy, qc = col + dx;
                if (qr >= 0 && qr < H && qc >= 0 && qc < W) {
                    const float* p = in + 3 * ((size_t)qc * H + qr);
                    const float l = dn_lum(p[0], p[1], p[2]);
                    s += l; s2 += l * l; count++;
                }
            }
        const float mean = s / (float)count, mean_sq = s2 / (float)count;
        variance[idx] = fmaxf(0.0f, mean_sq - mean * mean);
    }
}

struct DenoisePass { int W, H, step; float sigma_color, sigma_normal, sigma_depth; int use_variance; };

__global__ void __launch_bounds__(256) k_denoise_atrous(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ normals,
                                                         const float* __restrict__ depth, const float* __restrict__ variance, DenoisePass P) {     // :123-207
    const float K[5] = {1.0f / 16.0f, 1.0f / 4.0f, 3.0f / 8.0f, 1.0f / 4.0f, 1.0f / 16.0f};
    const int W = P.W, H = P.H;
    const uint32_t n = (uint32_t)W * (uint32_t)H;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int row = (int)(idx % (uint32_t)H), col = (int)(idx / (uint32_t)H);
        const float rp = in[3 * (size_t)idx], gp = in[3 * (size_t)idx + 1], bp = in[3 * (size_t)idx + 2];
        const float lum_p = dn_lum(rp, gp, bp);
        const float3 n_p = f3(normals[3 * (size_t)idx], normals[3 * (size_t)idx + 1], normals[3 * (size_t)idx + 2]);
        const float d_p = depth[idx];
        const float var_p = P.use_variance ? variance[idx] : 0.0f;
        const float es = var_p > 0.0f ? P.sigma_color * sqrtf(var_p) + 1.0e-4f : P.sigma_color;      // weight_color :83-89
        const float ds = P.sigma_depth * (float)P.step + 1.0e-4f;                                     // weight_depth :111-113
        float sr = 0.0f, sg = 0.0f, sb = 0.0f, sw = 0.0f;
        for (int dyi = 0; dyi < 5; dyi++)
            for (int dxi = 0; dxi < 5; dxi++) {
                int qr = row + (dyi - 2) * P.step, qc = col + (dxi - 2) * P.step;
                qr = min(max(qr, 0), H - 1); qc = min(max(qc, 0), W - 1);
                const size_t q = (size_t)qc * H + qr;
                const float rq = in[3 * q], gq = in[3 * q + 1], bq = in[3 * q + 2];
                const float lum_q = dn_lum(rq, gq, bq);
                const float3 n_q = f3(normals[3 * q], normals[3 * q + 1], normals[3 * q + 2]);
                const float w_spatial = K[dxi] * K[dyi];
                const float w_color = dm_expf(-fabsf(lum_p - lum_q) / es);
                const float dotv = dot3(n_p, n_q);
                const float w_norm = dm_powf(dotv > 0.0f ? dotv : (dotv == dotv ? 0.0f : dotv), P.sigma_normal);       // max(0, dot) ^ sigma; a NaN dot stays NaN as in Julia
                const float w_depth = dm_expf(-fabsf(d_p - depth[q]) / ds);
                const float w = w_spatial * w_color * w_norm * w_depth;
                sr += rq * w; sg += gq * w; sb += bq * w; sw += w;
            }
        float* o = out + 3 * (size_t)idx;
        if (sw > 1.0e-6f) { const float inv = 1.0f / sw; o[0] = sr * inv; o[1] = sg * inv; o[2] = sb * inv; }
        else { o[0] = rp; o[1] = gp; o[2] = bp; }      // also the NaN case (escaped pixel: |Inf - Inf|)
    }
}
