// hk_oracle.cpp — ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of Hikari.jl's VolPath hot path, exported with a C ABI that mirrors
// include/hikari_cuda.h (ok_* instead of hk_*) so tests can feed the SAME flattened scene to both.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// PARITY PINNING: the reference cannot run here (no Julia; Raycore.jl absent). The oracle is pinned by
// the golden vectors the reference does hold for this path (tests/test_oracle_golden.py):
// fresnel_dielectric zeros (test/materials.jl:3-4), the gray-RGB closed form (rgb2spec.jl:90-102,
// test/rgb2spec_gpu.jl:105-140), filter importance-sampling invariants (test/filter.jl), pbrt-v4's published
// MurmurHash64A / PCG32 / Sobol known answers, and the smoke bounds of test/volpath_integration.jl:99-114.
// Closest-hit ids and image values have NO upstream golden data => "parity unpinned" for those.
#include "ok_volpath.h"
#include <cstdio>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ok;

extern "C" {

struct OkContext { Scene s; };

int32_t ok_create(OkContext** out) { *out = new OkContext(); return 0; }
int32_t ok_destroy(OkContext* c) { delete c; return 0; }
int32_t ok_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int32_t ok_set_num_threads(int32_t n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    return 0;
}
int32_t ok_set_brute_force(OkContext* c, int32_t on) { c->s.brute_force = on != 0; return 0; }
// bounded sample of a frame (bench.py CPU legs): render only every `step`-th image row, starting at row `offset`
int32_t ok_set_row_subset(OkContext* c, int32_t step, int32_t offset) {
    if (step < 1 || offset < 0 || offset >= step) return -1;
    c->s.row_step = step; c->s.row_offset = offset; return 0;
}

int32_t ok_upload_tables(OkContext* c, const HkTables* t) {
    Scene& s = c->s;
    s.sobol.assign(t->sobol_matrices, t->sobol_matrices + 1024 * 52);
    s.cie_x.assign(t->cie_x, t->cie_x + 471); s.cie_y.assign(t->cie_y, t->cie_y + 471); s.cie_z.assign(t->cie_z, t->cie_z + 471);
    s.d65.assign(t->d65, t->d65 + 107);
    size_t r = (size_t)t->rgb2spec_res;
    s.rgb_scale.assign(t->rgb2spec_scale, t->rgb2spec_scale + r);
    s.rgb_coeffs.assign(t->rgb2spec_coeffs, t->rgb2spec_coeffs + 9 * r * r * r);
    s.T = Tables{s.sobol.data(), s.cie_x.data(), s.cie_y.data(), s.cie_z.data(), s.d65.data(), t->rgb2spec_res, s.rgb_scale.data(), s.rgb_coeffs.data()};
    return 0;
}
int32_t ok_upload_geometry(OkContext* c, const HkGeometry* g) {
    Scene& s = c->s;
    s.positions.assign(g->positions, g->positions + 3 * (size_t)g->n_verts);
    s.has_normals = g->normals != nullptr; s.has_tangents = g->tangents != nullptr; s.has_uvs = g->uvs != nullptr;
    if (g->normals) s.normals.assign(g->normals, g->normals + 3 * (size_t)g->n_verts); else s.normals.clear();
    if (g->tangents) s.tangents.assign(g->tangents, g->tangents + 3 * (size_t)g->n_verts); else s.tangents.clear();
    if (g->uvs) s.uvs.assign(g->uvs, g->uvs + 2 * (size_t)g->n_verts); else s.uvs.clear();
    s.indices.assign(g->indices, g->indices + 3 * (size_t)g->n_tris);
    if (g->n_instances > 0) {       // object-space meshes + instances (HkGeometry.instances): per-mesh accelerators, nothing flattened
        for (uint32_t i = 0; i < g->n_instances; i++) if (g->instances[i].mesh >= g->n_meshes) return -1;
        s.tri_meta.clear();
        s.accel = Accel();
        s.iaccel.build(s.positions.data(), s.indices.data(), g->meshes, g->n_meshes, g->instances, g->n_instances);
        return 0;
    }
    s.iaccel = InstancedAccel();
    s.tri_meta.assign(g->tri_meta, g->tri_meta + 3 * (size_t)g->n_tris);
    s.accel.build(s.positions.data(), s.indices.data(), g->n_tris);
    return 0;
}
int32_t ok_upload_spectra(OkContext* c, const HkSpectra* sp) {
    Scene& s = c->s;
    s.spec_offsets.assign(sp->offsets, sp->offsets + sp->n_spectra + 1);
    uint32_t n = sp->n_spectra ? sp->offsets[sp->n_spectra] : 0;
    s.spec_lambdas.assign(sp->lambdas, sp->lambdas + n); s.spec_values.assign(sp->values, sp->values + n);
    s.spectra = HkSpectra{s.spec_lambdas.data(), s.spec_values.data(), s.spec_offsets.data(), sp->n_spectra};
    return 0;
}
int32_t ok_upload_textures(OkContext* c, const HkTexture* t, uint32_t n) {
    TextureStore& S = c->s.textures;
    S.rgb.clear(); S.h.clear(); S.w.clear(); c->s.tex_alpha.assign(n, std::vector<float>());
    for (uint32_t i = 0; i < n; i++) {
        S.rgb.emplace_back(t[i].rgb, t[i].rgb + 3 * (size_t)t[i].h * t[i].w); S.h.push_back(t[i].h); S.w.push_back(t[i].w);
        if (t[i].alpha) c->s.tex_alpha[i].assign(t[i].alpha, t[i].alpha + (size_t)t[i].h * t[i].w);      // HkTexture::alpha, (h, w) column-major
    }
    return 0;
}
int32_t ok_upload_materials(OkContext* c, const HkMaterial* m, uint32_t nm, const HkMediumInterface* mi, uint32_t ni) {
    c->s.materials.assign(m, m + nm); c->s.interfaces.assign(mi, mi + ni);
    if (c->s.spec_offsets.empty()) { c->s.spec_offsets.assign(1, 0); c->s.spectra = HkSpectra{nullptr, nullptr, c->s.spec_offsets.data(), 0}; }
    return 0;
}
int32_t ok_update_material(OkContext* c, uint32_t index, const HkMaterial* m) {   // update_material!, scene.jl:109-112
    if (index < 1 || index > c->s.materials.size()) return -1;
    c->s.materials[index - 1] = *m;
    return 0;
}
int32_t ok_upload_envmaps(OkContext* c, const HkEnvMap* maps, uint32_t n) {
    Scene& s = c->s;
    s.envmaps.assign(maps, maps + n); s.env_store.clear(); s.env_store.reserve(6 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) {
        HkEnvMap& E = s.envmaps[i];
        size_t w = E.w, h = E.h, nu = E.nu, nv = E.nv;
        auto keep = [&](const float* p, size_t cnt) { s.env_store.emplace_back(p, p + cnt); return (const float*)s.env_store.back().data(); };
        E.rgb = keep(E.rgb, w * h * 3);
        E.conditional_func = keep(E.conditional_func, nu * nv);
        E.conditional_cdf = keep(E.conditional_cdf, (nu + 1) * nv);
        E.conditional_func_int = keep(E.conditional_func_int, nv);
        E.marginal_func = keep(E.marginal_func, nv);
        E.marginal_cdf = keep(E.marginal_cdf, nv + 1);
    }
    return 0;
}
int32_t ok_upload_lights(OkContext* c, const HkLight* l, uint32_t n, const HkLightSampler* sm) {
    Scene& s = c->s;
    s.lights.assign(l, l + n);
    s.lnodes.assign(sm->nodes, sm->nodes + sm->n_nodes);
    s.bit_trails.assign(sm->light_to_bit_trail, sm->light_to_bit_trail + n);
    s.inf_idx.assign(sm->infinite_light_indices, sm->infinite_light_indices + sm->n_infinite);
    s.sampler = HkLightSampler{s.lnodes.data(), sm->n_nodes, s.bit_trails.data(), s.inf_idx.data(), sm->n_infinite, sm->n_bvh_lights};
    return 0;
}
int32_t ok_upload_media(OkContext* c, const HkMedium* m, uint32_t n) {
    Scene& s = c->s;
    s.media.clear(); s.media.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        Medium& M = s.media[i]; M.h = m[i];
        if (m[i].type == HK_MEDIUM_GRID) {
            size_t cnt = (size_t)m[i].density_res[0] * m[i].density_res[1] * m[i].density_res[2];
            M.density.assign(m[i].density, m[i].density + cnt);
        }
        if (m[i].type != HK_MEDIUM_HOMOGENEOUS) {
            size_t cnt = (size_t)m[i].majorant_res[0] * m[i].majorant_res[1] * m[i].majorant_res[2];
            M.majorant.assign(m[i].majorant, m[i].majorant + cnt);
        }
        if (m[i].type == HK_MEDIUM_NANOVDB) M.nvdb.assign(m[i].nanovdb_buf, m[i].nanovdb_buf + m[i].nanovdb_bytes);
        if (m[i].type == HK_MEDIUM_RGBGRID) {
            size_t cnt = 3 * (size_t)m[i].density_res[0] * m[i].density_res[1] * m[i].density_res[2];
            if (m[i].rgb_sigma_a) M.rgb_a.assign(m[i].rgb_sigma_a, m[i].rgb_sigma_a + cnt);
            if (m[i].rgb_sigma_s) M.rgb_s.assign(m[i].rgb_sigma_s, m[i].rgb_sigma_s + cnt);
            if (m[i].rgb_Le) M.rgb_le.assign(m[i].rgb_Le, m[i].rgb_Le + cnt);
        }
        M.h.density = nullptr; M.h.majorant = nullptr; M.h.nanovdb_buf = nullptr; M.h.rgb_sigma_a = M.h.rgb_sigma_s = M.h.rgb_Le = nullptr;
    }
    return 0;
}
int32_t ok_set_camera(OkContext* c, const HkCamera* cam) { c->s.camera = *cam; return 0; }
int32_t ok_set_filter(OkContext* c, const HkFilter* f) {
    Scene& s = c->s;
    s.filter = *f;
    if (f->type >= 3) {
        size_t nx = f->nx, ny = f->ny;
        s.f_func.assign(f->func, f->func + nx * ny); s.f_mcdf.assign(f->marginal_cdf, f->marginal_cdf + ny + 1);
        s.f_mfunc.assign(f->marginal_func, f->marginal_func + ny); s.f_ccdf.assign(f->conditional_cdf, f->conditional_cdf + ny * (nx + 1));
        s.filter.func = s.f_func.data(); s.filter.marginal_cdf = s.f_mcdf.data(); s.filter.marginal_func = s.f_mfunc.data(); s.filter.conditional_cdf = s.f_ccdf.data();
    }
    return 0;
}
int32_t ok_set_params(OkContext* c, const HkRenderParams* p) { c->s.params = *p; c->s.alloc_film(); return 0; }
int32_t ok_clear(OkContext* c) {
    std::fill(c->s.pixel_rgb.begin(), c->s.pixel_rgb.end(), 0.0f);
    std::fill(c->s.pixel_weight_sum.begin(), c->s.pixel_weight_sum.end(), 0.0f);
    std::fill(c->s.aux_albedo.begin(), c->s.aux_albedo.end(), 0.0f); std::fill(c->s.aux_normal.begin(), c->s.aux_normal.end(), 0.0f);
    std::fill(c->s.aux_depth.begin(), c->s.aux_depth.end(), 0.0f);     // clear!(film), film.jl:343-346
    return 0;
}
int32_t ok_fill_aux_buffers(OkContext* c, int32_t has_infinite_lights) { c->s.fill_aux_buffers(has_infinite_lights != 0); return 0; }
int32_t ok_read_aux_buffers(OkContext* c, float* albedo, float* normal, float* depth) {
    const size_t n = (size_t)c->s.params.width * c->s.params.height;
    if (c->s.aux_depth.size() != n) return -1;
    if (albedo) std::memcpy(albedo, c->s.aux_albedo.data(), 12 * n);
    if (normal) std::memcpy(normal, c->s.aux_normal.data(), 12 * n);
    if (depth) std::memcpy(depth, c->s.aux_depth.data(), 4 * n);
    return 0;
}
int32_t ok_render_samples_strided(OkContext* c, int32_t first, int32_t stride, int32_t count) {
    for (int32_t i = 0; i < count; i++) c->s.render_sample(first + i * stride);
    return 0;
}
int32_t ok_render_samples(OkContext* c, int32_t first, int32_t count) { return ok_render_samples_strided(c, first, 1, count); }
// vp_finalize_film_kernel!, volpath.jl:384-417: framebuffer[py, px] (H, W) column-major
int32_t ok_read_film(OkContext* c, float* out) {
    Scene& s = c->s;
    const int W = s.params.width, H = s.params.height;
    for (int64_t p = 0; p < (int64_t)W * H; p++) {
        int px = (int)(p % W), py = (int)(p / W);
        float ws = s.pixel_weight_sum[p];
        float r = 0, g = 0, b = 0;
        if (ws > 0.0f) { float inv = 1.0f / ws; r = s.pixel_rgb[3 * p] * inv; g = s.pixel_rgb[3 * p + 1] * inv; b = s.pixel_rgb[3 * p + 2] * inv; }
        float* o = out + ((size_t)px * H + py) * 3;
        o[0] = r; o[1] = g; o[2] = b;
    }
    return 0;
}
// postprocess!, src/postprocess.jl:55-182 (tone maps), :187-230 (per-pixel kernel), fused with the film read-out like
// hk_postprocess.  Plain scalar C++ restatement; libm powf for the gamma.
static float pp_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static float pp_uncharted2(float x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
static float pp_filmic(float x) { x = std::max(0.0f, x - 0.004f); return (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f); }
static float pp_aces(float x) { return pp_clamp01((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f)); }
int32_t ok_postprocess(OkContext* c, const HkPostprocess* P, float* out) {
    Scene& s = c->s;
    const int W = s.params.width, H = s.params.height;
    for (int64_t p = 0; p < (int64_t)W * H; p++) {
        int px = (int)(p % W), py = (int)(p / W);
        float ws = s.pixel_weight_sum[p];
        float r = 0, g = 0, b = 0;
        if (ws > 0.0f) { float inv = 1.0f / ws; r = s.pixel_rgb[3 * p] * inv; g = s.pixel_rgb[3 * p + 1] * inv; b = s.pixel_rgb[3 * p + 2] * inv; }
        r *= P->exposure; g *= P->exposure; b *= P->exposure;
        if (P->apply_wb) {
            const float* m = P->wb;
            float ro = m[0] * r + m[1] * g + m[2] * b, go = m[3] * r + m[4] * g + m[5] * b, bo = m[6] * r + m[7] * g + m[8] * b;
            r = std::max(0.0f, ro); g = std::max(0.0f, go); b = std::max(0.0f, bo);
        }
        r *= P->imaging_ratio; g *= P->imaging_ratio; b *= P->imaging_ratio;
        switch (P->tonemap_mode) {
            case HK_TONEMAP_REINHARD: {
                float lum = 0.2126f * r + 0.7152f * g + 0.0722f * b, sc = lum > 0.0f ? 1.0f / (1.0f + lum) : 1.0f;
                r = pp_clamp01(r * sc); g = pp_clamp01(g * sc); b = pp_clamp01(b * sc); break; }
            case HK_TONEMAP_REINHARD_EXT: {
                float lum = 0.2126f * r + 0.7152f * g + 0.0722f * b, lw2 = P->white_point * P->white_point;
                float sc = lum > 0.0f ? (1.0f + lum / lw2) / (1.0f + lum) : 1.0f;
                r = pp_clamp01(r * sc); g = pp_clamp01(g * sc); b = pp_clamp01(b * sc); break; }
            case HK_TONEMAP_ACES: r = pp_aces(r); g = pp_aces(g); b = pp_aces(b); break;
            case HK_TONEMAP_UNCHARTED2: {
                float wsc = 1.0f / pp_uncharted2(11.2f);
                r = pp_clamp01(pp_uncharted2(r * 2.0f) * wsc); g = pp_clamp01(pp_uncharted2(g * 2.0f) * wsc); b = pp_clamp01(pp_uncharted2(b * 2.0f) * wsc); break; }
            case HK_TONEMAP_FILMIC: r = pp_filmic(r); g = pp_filmic(g); b = pp_filmic(b); break;
            default: r = pp_clamp01(r); g = pp_clamp01(g); b = pp_clamp01(b); break;
        }
        if (P->apply_gamma) { r = dm_powf(r, P->inv_gamma); g = dm_powf(g, P->inv_gamma); b = dm_powf(b, P->inv_gamma); }
        if (P->mask_escaped) {     // AA depth mask, postprocess.jl:220-245: (row, col) = (py + 1, px + 1) of the (H, W) array, depth row flipped
            if (s.aux_depth.size() != (size_t)W * H) return -1;
            const int d_row = H - (py + 1) + 1, col = px + 1;
            int escaped = 0, total = 0;
            for (int dr = -1; dr <= 1; dr++)
                for (int dc = -1; dc <= 1; dc++) {
                    const int nr = d_row + dr, nc = col + dc;
                    if (nr >= 1 && nr <= H && nc >= 1 && nc <= W) { escaped += std::isinf(s.aux_depth[(size_t)(nc - 1) * H + (nr - 1)]) ? 1 : 0; total++; }
                }
            const float alpha = (float)escaped / (float)total;
            r = r * (1.0f - alpha) + P->background[0] * alpha; g = g * (1.0f - alpha) + P->background[1] * alpha; b = b * (1.0f - alpha) + P->background[2] * alpha;
        }
        float* o = out + ((size_t)px * H + py) * 3;
        o[0] = r; o[1] = g; o[2] = b;
    }
    return 0;
}
int32_t ok_test_set_aux_depth(OkContext* c, const float* depth) {      // test hook: a synthetic film.depth ((H, W) column-major)
    const size_t n = (size_t)c->s.params.width * c->s.params.height;
    if (c->s.aux_depth.size() != n) return -1;
    std::memcpy(c->s.aux_depth.data(), depth, 4 * n);
    return 0;
}
// denoise!(film; config), src/denoise.jl:301-372 (weights :66-114, a-trous pass :123-207, variance :216-258); images (H, W) column-major
static float dn_lum(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }
static float dn_max0(float x) { return x > 0.0f ? x : (x == x ? 0.0f : x); }      // Julia max(0f0, NaN) = NaN
int32_t ok_denoise(OkContext* c, const HkDenoiseConfig* cfg, float* out_pp, float* out_fb) {
    Scene& s = c->s;
    const int W = s.params.width, H = s.params.height;
    const size_t n = (size_t)W * H;
    if (s.aux_depth.size() != n || cfg->iterations < 0) return -1;
    std::vector<float> a(3 * n), b(3 * n), var(n, 0.0f);
    for (size_t p = 0; p < n; p++) {      // film.framebuffer (vp_finalize_film_kernel!)
        const size_t px = p % W, py = p / W;
        const float ws = s.pixel_weight_sum[p];
        float* o = a.data() + 3 * (px * H + py);
        if (ws > 0.0f) { const float inv = 1.0f / ws; o[0] = s.pixel_rgb[3 * p] * inv; o[1] = s.pixel_rgb[3 * p + 1] * inv; o[2] = s.pixel_rgb[3 * p + 2] * inv; }
        else o[0] = o[1] = o[2] = 0.0f;
    }
    if (cfg->use_variance) {
        for (size_t idx = 0; idx < n; idx++) {
            const int row = (int)(idx % H), col = (int)(idx / H);
            float sl = 0.0f, sl2 = 0.0f; int count = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int qr = row + dy, qc = col + dx;
                    if (qr >= 0 && qr < H && qc >= 0 && qc < W) {
                        const float* p = a.data() + 3 * ((size_t)qc * H + qr);
                        const float l = dn_lum(p[0], p[1], p[2]);
                        sl += l; sl2 += l * l; count++;
                    }
                }
            const float mean = sl / (float)count, mean_sq = sl2 / (float)count;
            var[idx] = std::max(0.0f, mean_sq - mean * mean);
        }
    }
    const float K[5] = {1.0f / 16.0f, 1.0f / 4.0f, 3.0f / 8.0f, 1.0f / 4.0f, 1.0f / 16.0f};
    for (int it = 1; it <= cfg->iterations; it++) {
        const int step = 1 << (it - 1);
        const std::vector<float>& in = (it % 2 == 1) ? a : b;
        std::vector<float>& out = (it % 2 == 1) ? b : a;
        #pragma omp parallel for schedule(static)
        for (int64_t idx = 0; idx < (int64_t)n; idx++) {
            const int row = (int)(idx % H), col = (int)(idx / H);
            const float rp = in[3 * idx], gp = in[3 * idx + 1], bp = in[3 * idx + 2];
            const float lum_p = dn_lum(rp, gp, bp);
            const float* np_ = s.aux_normal.data() + 3 * idx;
            const float d_p = s.aux_depth[idx];
            const float var_p = cfg->use_variance ? var[idx] : 0.0f;
            float sr = 0.0f, sg = 0.0f, sb = 0.0f, sw = 0.0f;
            for (int dyi = 0; dyi < 5; dyi++)
                for (int dxi = 0; dxi < 5; dxi++) {
                    int qr = row + (dyi - 2) * step, qc = col + (dxi - 2) * step;
                    qr = std::min(std::max(qr, 0), H - 1); qc = std::min(std::max(qc, 0), W - 1);
                    const size_t q = (size_t)qc * H + qr;
                    const float rq = in[3 * q], gq = in[3 * q + 1], bq = in[3 * q + 2];
                    const float lum_q = dn_lum(rq, gq, bq);
                    const float* nq = s.aux_normal.data() + 3 * q;
                    const float w_spatial = K[dxi] * K[dyi];
                    const float es = var_p > 0.0f ? cfg->sigma_color * std::sqrt(var_p) + 1.0e-4f : cfg->sigma_color;
                    const float w_color = dm_expf(-std::fabs(lum_p - lum_q) / es);
                    const float dotv = (np_[0] * nq[0] + np_[1] * nq[1]) + np_[2] * nq[2];
                    const float w_norm = dm_powf(dn_max0(dotv), cfg->sigma_normal);
                    const float w_depth = dm_expf(-std::fabs(d_p - s.aux_depth[q]) / (cfg->sigma_depth * (float)step + 1.0e-4f));
                    const float w = w_spatial * w_color * w_norm * w_depth;
                    sr += rq * w; sg += gq * w; sb += bq * w; sw += w;
                }
            float* o = out.data() + 3 * idx;
            if (sw > 1.0e-6f) { const float inv = 1.0f / sw; o[0] = sr * inv; o[1] = sg * inv; o[2] = sb * inv; }
            else { o[0] = rp; o[1] = gp; o[2] = bp; }
        }
    }
    std::memcpy(out_pp, (cfg->iterations % 2 == 1 ? b : a).data(), 12 * n);
    if (out_fb && cfg->iterations >= 2) std::memcpy(out_fb, a.data(), 12 * n);
    return 0;
}
int32_t ok_read_accum(OkContext* c, float* rgb, float* w) {
    std::memcpy(rgb, c->s.pixel_rgb.data(), c->s.pixel_rgb.size() * 4);
    std::memcpy(w, c->s.pixel_weight_sum.data(), c->s.pixel_weight_sum.size() * 4);
    return 0;
}
// last sample pass' spectral buffer + per-pixel wavelengths (for stage-level parity checks)
int32_t ok_write_accum(OkContext* c, const float* rgb, const float* w) {    // test hook: same film state on both paths
    const size_t n = (size_t)c->s.params.width * c->s.params.height;
    std::memcpy(c->s.pixel_rgb.data(), rgb, 12 * n); std::memcpy(c->s.pixel_weight_sum.data(), w, 4 * n);
    return 0;
}
int32_t ok_read_pixel_L(OkContext* c, float* L, float* lambda, float* pdf, float* fw) {
    Scene& s = c->s;
    std::memcpy(L, s.pixel_L.data(), s.pixel_L.size() * 4); std::memcpy(lambda, s.wavelengths.data(), s.wavelengths.size() * 4);
    std::memcpy(pdf, s.pdfs.data(), s.pdfs.size() * 4); std::memcpy(fw, s.filter_weight.data(), s.filter_weight.size() * 4);
    return 0;
}
uint64_t ok_rays_traced(OkContext* c) { return c->s.rays_traced; }

// rays [n][8] -> hits [n][4] (t, prim as u32 bits 1-based / 0 = miss, b1, b2); mode 0 = BVH2, 1 = brute force
int32_t ok_trace_closest(OkContext* c, const float* rays, uint64_t n, float* hits, int32_t mode) {
    const Scene& s = c->s;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* r = rays + 8 * i;
        V3 o(r[0], r[1], r[2]), d(r[3], r[4], r[5]);
        Hit h = s.iaccel.enabled() ? s.iaccel.closest_hit(o, d, r[6], mode != 0) : (mode ? s.accel.closest_hit_brute(o, d, r[6]) : s.accel.closest_hit_bvh(o, d, r[6]));
        float* out = hits + 4 * i;
        uint32_t prim = h.hit ? h.prim + 1 : 0;
        out[0] = h.hit ? h.t : r[6]; std::memcpy(out + 1, &prim, 4); out[2] = h.hit ? h.b1 : 0.0f; out[3] = h.hit ? h.b2 : 0.0f;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// batch entry points for per-function parity tests (each mirrors an hk_test_* kernel)
// ---------------------------------------------------------------------------------------------
// out[i] = zsobol 1d / 2d samples for (px,py,sample_idx,dim) quadruples; 2d writes 2 floats
// fn: 0 expf 1 logf 2 sinf 3 cosf 4 coshf 5 atanhf 6 powf(x, y) 7 log1pf  (hk_detmath.h; the CUDA library runs the same source)
int32_t ok_test_detmath(int32_t fn, const float* x, const float* y, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; i++) {
        const float a = x[i], b = y ? y[i] : 0.0f;
        out[i] = fn == 0 ? dm_expf(a) : fn == 1 ? dm_logf(a) : fn == 2 ? dm_sinf(a) : fn == 3 ? dm_cosf(a) : fn == 4 ? dm_coshf(a) : fn == 5 ? dm_atanhf(a) : fn == 6 ? dm_powf(a, b) : dm_log1pf(a);
    }
    return 0;
}
int32_t ok_test_sobol(OkContext* c, const int32_t* q, uint64_t n, int32_t log2_spp, int32_t nb4, uint32_t seed, float* out1d, float* out2d) {
    SobolRNG r{c->s.T.sobol, log2_spp, nb4, seed, 0};
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const int32_t* e = q + 4 * i;
        out1d[i] = zsobol_1d(r, e[0], e[1], e[2], e[3]);
        V2 v = zsobol_2d(r, e[0], e[1], e[2], e[3]);
        out2d[2 * i] = v.x; out2d[2 * i + 1] = v.y;
    }
    return 0;
}
int32_t ok_test_hashes(const float* v3, uint64_t n, uint64_t* out_hash, uint64_t* out_mix, float* out_pcg) {
    for (uint64_t i = 0; i < n; i++) {
        V3 v(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2]);
        out_hash[i] = pbrt_hash(v);
        out_mix[i] = mix_bits(out_hash[i]);
        PCG32 r = pcg32_init(out_hash[i], out_mix[i]);
        out_pcg[2 * i] = pcg32_f32(r); out_pcg[2 * i + 1] = pcg32_f32(r);
    }
    return 0;
}
uint64_t ok_murmur64a(const uint8_t* data, uint64_t n, uint64_t seed) { return murmur_hash_64a(data, (size_t)n, seed); }
int32_t ok_pcg32_stream(uint64_t seq, uint64_t seed, uint32_t* out, int32_t n) {
    PCG32 r = pcg32_init(seq, seed);
    for (int i = 0; i < n; i++) out[i] = pcg32_u32(r);
    return 0;
}
int32_t ok_sobol_raw(OkContext* c, int64_t a, int32_t dim, uint32_t* out_bits) {   // unscrambled Sobol integer
    uint32_t v = 0; const uint32_t* M = c->s.T.sobol;
    for (int bit = 0; bit < 52; bit++) if ((a >> bit) & 1) v ^= M[dim * 52 + bit];
    *out_bits = v; return 0;
}
// wavelengths: u[n] -> lambda[n][4], pdf[n][4]
int32_t ok_test_wavelengths(const float* u, uint64_t n, float* lambda, float* pdf) {
    for (uint64_t i = 0; i < n; i++) {
        Wavelengths w = sample_wavelengths_visible(u[i]);
        for (int k = 0; k < 4; k++) { lambda[4 * i + k] = w.lambda[k]; pdf[4 * i + k] = w.pdf[k]; }
    }
    return 0;
}
// kind: 0 uplift_rgb, 1 unbounded, 2 illuminant; rgb[n][3], lambda[n][4] -> out[n][4]; coeffs -> poly[n][3]
int32_t ok_test_uplift(OkContext* c, int32_t kind, const float* rgb, const float* lambda, uint64_t n, float* out, float* poly) {
    const Tables& T = c->s.T;
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Wavelengths w; for (int k = 0; k < 4; k++) { w.lambda[k] = lambda[4 * i + k]; w.pdf[k] = 1.0f; }
        Spec s = kind == 0 ? uplift_rgb(T, rgb + 3 * i, w) : (kind == 1 ? uplift_rgb_unbounded(T, rgb + 3 * i, w) : uplift_rgb_illuminant(T, rgb + 3 * i, w));
        for (int k = 0; k < 4; k++) out[4 * i + k] = s.v[k];
        Poly p = rgb_to_spectrum(T, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        poly[3 * i] = p.c0; poly[3 * i + 1] = p.c1; poly[3 * i + 2] = p.c2;
    }
    return 0;
}
// L[n][4], lambda[n][4], pdf[n][4] -> rgb[n][3] (spectral_to_xyz -> xyz_to_linear_srgb, no clamp)
int32_t ok_test_spectral_to_rgb(OkContext* c, const float* L, const float* lambda, const float* pdf, uint64_t n, float* xyz, float* rgb) {
    for (uint64_t i = 0; i < n; i++) {
        Wavelengths w; for (int k = 0; k < 4; k++) { w.lambda[k] = lambda[4 * i + k]; w.pdf[k] = pdf[4 * i + k]; }
        V3 x = spectral_to_xyz(c->s.T, Spec(L[4 * i], L[4 * i + 1], L[4 * i + 2], L[4 * i + 3]), w);
        V3 r = xyz_to_linear_srgb(x);
        xyz[3 * i] = x.x; xyz[3 * i + 1] = x.y; xyz[3 * i + 2] = x.z; rgb[3 * i] = r.x; rgb[3 * i + 1] = r.y; rgb[3 * i + 2] = r.z;
    }
    return 0;
}
int32_t ok_test_filter(OkContext* c, const float* u, uint64_t n, float* out /*[n][3] px py w*/) {
    for (uint64_t i = 0; i < n; i++) {
        FilterSample f = filter_sample(c->s.filter, V2(u[2 * i], u[2 * i + 1]));
        out[3 * i] = f.p.x; out[3 * i + 1] = f.p.y; out[3 * i + 2] = f.weight;
    }
    return 0;
}
// camera rays for a full sample pass: out[n_pixels][8] o,d,lambda0,filter_weight
int32_t ok_test_camera_rays(OkContext* c, int32_t sample_idx, float* out) {
    Scene& s = c->s;
    const int W = s.params.width, H = s.params.height;
    SobolRNG rng{s.T.sobol, s.params.sobol_log2_spp, s.params.sobol_n_base4_digits, s.params.sampler_seed, W};
    #pragma omp parallel for
    for (int64_t p = 0; p < (int64_t)W * H; p++) {
        int x = (int)(p % W) + 1, y = (int)(p / W) + 1;
        float wu = zsobol_1d(rng, x, y, sample_idx, 1);
        V2 j = zsobol_2d(rng, x, y, sample_idx, 3);
        float tu = zsobol_1d(rng, x, y, sample_idx, 4);
        V2 lens = zsobol_2d(rng, x, y, sample_idx, 6);
        FilterSample fs = filter_sample(s.filter, j);
        Wavelengths lam = sample_wavelengths_visible(wu);
        V2 pf((float)x + 0.5f + fs.p.x, (float)H - (float)y + 1.0f + 0.5f + fs.p.y);
        Ray r = camera_generate_ray(s.camera, pf, lens, tu);
        float* o = out + 8 * p;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z; o[6] = lam.lambda[0]; o[7] = fs.weight;
    }
    return 0;
}
// BSDF batch: per item in[16] = wo.xyz, n.xyz, lambda[4], u.xy, uc, regularize, wi.xyz(for eval)
// out[16] = sample: wi.xyz, f[4], pdf, is_specular, eta_scale ; eval: f[4], pdf
int32_t ok_test_bsdf(OkContext* c, uint32_t material_idx, const float* in, uint64_t n, float* out) {
    const Scene& s = c->s;
    MatCtx MC = s.matctx();
    const HkMaterial& m = s.materials[material_idx - 1];
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* e = in + 17 * i; float* o = out + 16 * i;
        V3 wo(e[0], e[1], e[2]), nn(e[3], e[4], e[5]);
        Wavelengths w; for (int k = 0; k < 4; k++) { w.lambda[k] = e[6 + k]; w.pdf[k] = 1.0f; }
        BSDFSample bs = sample_material(MC, m, wo, nn, w, V2(e[10], e[11]), e[12], e[13] != 0.0f);
        o[0] = bs.wi.x; o[1] = bs.wi.y; o[2] = bs.wi.z; for (int k = 0; k < 4; k++) o[3 + k] = bs.f.v[k];
        o[7] = bs.pdf; o[8] = bs.is_specular ? 1.0f : 0.0f; o[9] = bs.eta_scale;
        BSDFEval be = eval_material(MC, m, wo, V3(e[14], e[15], e[16]), nn, w);
        for (int k = 0; k < 4; k++) o[10 + k] = be.f.v[k];
        o[14] = be.pdf; o[15] = 0.0f;
    }
    return 0;
}
// light batch: in[10] = p.xyz, n.xyz, lambda0 (4 wavelengths derived as in sample_wavelengths_visible(u)), uc, u.xy
// out[16] = light_idx, pmf, Li[4], wi.xyz, pdf, p_light.xyz, is_delta, pmf_replay, 0
int32_t ok_test_lights(OkContext* c, const float* in, uint64_t n, float* out) {
    const Scene& s = c->s;
    LightCtx LC = s.lightctx();
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* e = in + 10 * i; float* o = out + 16 * i;
        for (int k = 0; k < 16; k++) o[k] = 0.0f;
        V3 p(e[0], e[1], e[2]), nn(e[3], e[4], e[5]);
        Wavelengths w = sample_wavelengths_visible(e[6]);
        float pmf; int32_t li = bvh_sample_light(s.sampler, p, nn, e[7], pmf);
        o[0] = (float)li; o[1] = pmf;
        if (li >= 1 && li <= (int32_t)s.lights.size()) {
            LightSample ls = sample_light(LC, s.lights[li - 1], p, w, V2(e[8], e[9]));
            for (int k = 0; k < 4; k++) o[2 + k] = ls.Li.v[k];
            o[6] = ls.wi.x; o[7] = ls.wi.y; o[8] = ls.wi.z; o[9] = ls.pdf; o[10] = ls.p_light.x; o[11] = ls.p_light.y; o[12] = ls.p_light.z;
            o[13] = ls.is_delta ? 1.0f : 0.0f;
            o[14] = bvh_pmf(s.sampler, p, nn, li);
        }
    }
    return 0;
}
// escaped-ray batch: in[4] = d.xyz, lambda_u -> out[5] = Le[4], env pdf
int32_t ok_test_escaped(OkContext* c, const float* in, uint64_t n, float* out) {
    const Scene& s = c->s; LightCtx LC = s.lightctx();
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* e = in + 4 * i;
        V3 d(e[0], e[1], e[2]);
        Wavelengths w = sample_wavelengths_visible(e[3]);
        Spec Le = evaluate_escaped_ray(LC, d, w);
        for (int k = 0; k < 4; k++) out[5 * i + k] = Le.v[k];
        out[5 * i + 4] = compute_env_light_pdf(LC, d);
    }
    return 0;
}
// medium batch: in[8] = o.xyz, d.xyz, t_max, lambda_u ; out[16] = event, beta[4], r_u[4], r_l[4], p.xyz (scatter)
int32_t ok_test_delta_tracking(OkContext* c, uint32_t medium, const float* in, uint64_t n, float* out) {
    const Scene& s = c->s; MediaCtx MC = s.mediactx();
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* e = in + 8 * i; float* o = out + 16 * i;
        Wavelengths w = sample_wavelengths_visible(e[7]);
        DeltaResult r = delta_track(MC, medium, V3(e[0], e[1], e[2]), V3(e[3], e[4], e[5]), e[6], w, Spec(1.0f), Spec(1.0f), Spec(1.0f), 0, 1 << 30, nullptr);
        o[0] = (float)(int)r.event;
        for (int k = 0; k < 4; k++) { o[1 + k] = r.beta.v[k]; o[5 + k] = r.r_u.v[k]; o[9 + k] = r.r_l.v[k]; }
        o[13] = r.event == DeltaResult::SCATTER ? r.p.x : 0.0f; o[14] = r.event == DeltaResult::SCATTER ? r.p.y : 0.0f; o[15] = r.event == DeltaResult::SCATTER ? r.p.z : 0.0f;
    }
    return 0;
}
// density batch: p[n][3] -> sigma-scale density (Grid / NanoVDB) for medium idx
int32_t ok_test_density(OkContext* c, uint32_t medium, const float* p, uint64_t n, float* out) {
    const Scene& s = c->s; const Medium& m = s.media[medium - 1];
    for (uint64_t i = 0; i < n; i++) {
        V3 q(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        out[i] = m.h.type == HK_MEDIUM_GRID ? sample_grid_density(m, affine_point(m.h.medium_from_render, q)) : (m.h.type == HK_MEDIUM_NANOVDB ? sample_nanovdb_density(m, q) : 1.0f);
    }
    return 0;
}
// ratio-tracking batch: in[8] = o.xyz, d.xyz, t_max, lambda_u -> out[12] = T[4], r_u[4], r_l[4]
int32_t ok_test_ratio_tracking(OkContext* c, uint32_t medium, const float* in, uint64_t n, float* out) {
    const Scene& s = c->s; MediaCtx MC = s.mediactx();
    #pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float* e = in + 8 * i; float* o = out + 12 * i;
        Wavelengths w = sample_wavelengths_visible(e[7]);
        Spec T, ru, rl;
        transmittance_ratio_tracking(MC, medium, V3(e[0], e[1], e[2]), V3(e[3], e[4], e[5]), e[6], w, T, ru, rl);
        for (int k = 0; k < 4; k++) { o[k] = T.v[k]; o[4 + k] = ru.v[k]; o[8 + k] = rl.v[k]; }
    }
    return 0;
}
float ok_fresnel_dielectric(float c, float eta) { return fresnel_dielectric(c, eta); }
float ok_fr_complex(float c, float eta, float k) { return fr_complex(c, eta, k); }
// mix_hash_float, src/materials/mix-material.jl:114-158
float ok_mix_hash_float(const float* p, const float* wo, uint32_t type1, uint32_t vec1, uint32_t type2, uint32_t vec2) {
    return mix_hash_float(V3{p[0], p[1], p[2]}, V3{wo[0], wo[1], wo[2]}, type1, vec1, type2, vec2);
}

}  // extern "C"
