// ok_media.h — ORACLE (test infrastructure, NOT product code).
// Participating media: HG phase, majorant DDA, Homogeneous / Grid / NanoVDB density lookup,
// delta tracking, ratio tracking, medium NEE + phase sampling.
// Restates src/integrators/volpath/{media,nanovdb,delta-tracking,medium-scatter,intersection}.jl.
#pragma once
#include "ok_core.h"
#include "ok_spectral.h"
#include "ok_lights.h"
#include <vector>

namespace ok {

struct Medium {   // deep copy of an HkMedium
    HkMedium h;
    std::vector<float> density, majorant;
    std::vector<uint8_t> nvdb;
    std::vector<float> rgb_a, rgb_s, rgb_le;     // RGBGridMedium grids, [nz][ny][nx][3]; empty = absent
};
struct MediaCtx { const Tables* T; const Medium* media; uint32_t n; };

// media.jl:28-32
inline float hg_p(float g, float c) {
    float g2 = g * g;
    float denom = 1.0f + g2 - 2.0f * g * c;
    return (1.0f - g2) / (4.0f * PI_F * denom * std::sqrt(denom));
}
// media.jl:42-74
inline V3 sample_hg(float g, V3 wo, V2 u, float& pdf) {
    float c;
    if (std::fabs(g) < 1.0e-3f) c = 1.0f - 2.0f * u.x;
    else {
        float g2 = g * g;
        float sq = (1.0f - g2) / (1.0f - g + 2.0f * g * u.x);
        c = clampf((1.0f + g2 - sq * sq) / (2.0f * g), -1.0f, 1.0f);
    }
    float s = std::sqrt(std::max(0.0f, 1.0f - c * c));
    float phi = 2.0f * PI_F * u.y;
    V3 t1, t2; coordinate_system(-wo, t1, t2);
    V3 wi = normalize(s * dm_cosf(phi) * t1 + s * dm_sinf(phi) * t2 + c * (-wo));
    pdf = hg_p(g, c);
    return wi;
}

struct MediumProps { Spec sigma_a, sigma_s, Le; float g; };
struct MajSeg { float t_min, t_max; Spec sigma_maj; };
struct MajIter {   // RayMajorantIterator, media.jl:497-510
    int mode; Spec sigma_t; float t_min, t_max; bool hom_called;
    const float* grid; int res[3];
    float next_t[3], delta_t[3]; int step[3], limit[3], voxel[3];
};
inline MajIter majiter_invalid() { MajIter it{}; it.mode = 0; it.t_min = INF_F; it.t_max = -INF_F; it.hom_called = true; return it; }

inline float jmax(float a, float b) { return (a != a || b != b) ? NAN : std::max(a, b); }
inline float jmin(float a, float b) { return (a != a || b != b) ? NAN : std::min(a, b); }
// media.jl:1704-1740
inline void ray_bounds_intersect(V3 o, V3 d, const float* bmin, const float* bmax, float& t_enter, float& t_exit) {
    float t0[3], t1[3];
    for (int k = 0; k < 3; k++) {
        float inv = std::fabs(d[k]) > 1.0e-10f ? 1.0f / d[k] : (d[k] >= 0 ? INF_F : -INF_F);
        t0[k] = (bmin[k] - o[k]) * inv; t1[k] = (bmax[k] - o[k]) * inv;
        if (t0[k] > t1[k]) std::swap(t0[k], t1[k]);
    }
    t_enter = jmax(jmax(t0[0], t0[1]), t0[2]);
    t_exit = jmin(jmin(t1[0], t1[1]), t1[2]);
}
// media.jl:275-391
inline MajIter create_dda_iterator(const float* grid, const int* res, const float* bmin, const float* bmax, V3 o, V3 d, float t_min, float t_max, const Spec& sigma_t) {
    MajIter it{};
    it.sigma_t = sigma_t; it.t_min = t_min; it.t_max = t_max; it.grid = grid; it.hom_called = false;
    for (int k = 0; k < 3; k++) {
        it.res[k] = res[k];
        float diag = bmax[k] - bmin[k];
        float go = (o[k] - bmin[k]) / diag;
        float inv_diag = std::fabs(diag) > 1.0e-10f ? 1.0f / diag : 0.0f;
        float gd = d[k] * inv_diag;
        float gi = go + gd * t_min;
        int r = res[k];
        int vox = clampi(floor_int32(gi * (float)r), 0, r - 1);
        it.delta_t[k] = std::fabs(gd) > 1.0e-10f ? 1.0f / (std::fabs(gd) * (float)r) : INF_F;
        if (gd >= 0.0f) {
            float nvp = (float)(vox + 1) / (float)r;
            it.next_t[k] = gd > 1.0e-10f ? t_min + (nvp - gi) / gd : INF_F;
            it.step[k] = 1; it.limit[k] = r;
        } else {
            float nvp = (float)vox / (float)r;
            it.next_t[k] = gd < -1.0e-10f ? t_min + (nvp - gi) / gd : INF_F;
            it.step[k] = -1; it.limit[k] = -1;
        }
        it.voxel[k] = vox;
    }
    it.mode = (t_min >= t_max) ? 0 : 2;
    return it;
}
// media.jl:625-729
inline bool ray_majorant_next(MajIter& it, MajSeg& seg) {
    if (it.mode == 0) return false;
    if (it.mode == 1) {
        if (it.hom_called || it.t_min >= it.t_max) { it.mode = 0; it.hom_called = true; return false; }
        seg = MajSeg{it.t_min, it.t_max, it.sigma_t};
        it.hom_called = true;
        return true;
    }
    if (it.t_min >= it.t_max) { it.mode = 0; it.t_min = INF_F; it.t_max = -INF_F; it.hom_called = true; return false; }
    int axis = (it.next_t[0] < it.next_t[1]) ? ((it.next_t[0] < it.next_t[2]) ? 0 : 2) : ((it.next_t[1] < it.next_t[2]) ? 1 : 2);
    float seg_t_max = std::min(it.next_t[axis], it.t_max);
    float rho = it.grid[it.voxel[0] + it.res[0] * (it.voxel[1] + it.res[1] * it.voxel[2])];
    seg = MajSeg{it.t_min, seg_t_max, it.sigma_t * rho};
    it.t_min = seg_t_max;
    it.voxel[axis] += it.step[axis];
    it.next_t[axis] += it.delta_t[axis];
    if (it.voxel[0] == it.limit[0] || it.voxel[1] == it.limit[1] || it.voxel[2] == it.limit[2]) { it.mode = 0; it.t_min = it.t_max; }
    return true;
}

// ---- NanoVDB, nanovdb.jl:246-469 (byte offsets are 0-based here) ------------------------------------
template <class T> inline T rd(const uint8_t* buf, uint64_t off) { T v; std::memcpy(&v, buf + off, sizeof(T)); return v; }
inline bool bitmask_on(const uint8_t* buf, uint64_t mask_off, uint32_t n) { return ((buf[mask_off + (n >> 3)] >> (n & 7)) & 1) != 0; }
inline float nanovdb_get_value(const Medium& m, int32_t x, int32_t y, int32_t z) {
    const uint8_t* buf = m.nvdb.data();
    const HkMedium& h = m.h;
    uint32_t xu = (uint32_t)x, yu = (uint32_t)y, zu = (uint32_t)z;
    uint64_t key = (uint64_t)((zu >> 12) & 0x1fffff) | ((uint64_t)((yu >> 12) & 0x1fffff) << 21) | ((uint64_t)((xu >> 12) & 0x1fffff) << 42);
    uint64_t root = h.nanovdb_root_offset;
    uint64_t tile_base = root + 64, tile_off = 0; bool found = false;
    for (int32_t i = 0; i < h.nanovdb_root_tiles; i++) {
        uint64_t to = tile_base + (uint64_t)i * 32;
        if (rd<uint64_t>(buf, to) == key) { found = true; tile_off = to; break; }
    }
    if (!found) return rd<float>(buf, root + 28);
    int64_t child = rd<int64_t>(buf, tile_off + 8);
    if (child == 0) return rd<float>(buf, tile_off + 20);
    uint64_t upper = root + child;
    uint32_t nu = (((xu >> 7) & 31) << 10) | (((yu >> 7) & 31) << 5) | ((zu >> 7) & 31);
    if (!bitmask_on(buf, upper + 4128, nu)) return rd<float>(buf, upper + 8256 + (uint64_t)nu * 8);
    uint64_t lower = upper + rd<int64_t>(buf, upper + 8256 + (uint64_t)nu * 8);
    uint32_t nl = (((xu >> 3) & 15) << 8) | (((yu >> 3) & 15) << 4) | ((zu >> 3) & 15);
    if (!bitmask_on(buf, lower + 544, nl)) return rd<float>(buf, lower + 1088 + (uint64_t)nl * 8);
    uint64_t leaf = lower + rd<int64_t>(buf, lower + 1088 + (uint64_t)nl * 8);
    uint32_t nf = ((uint32_t)(x & 7) << 6) | ((uint32_t)(y & 7) << 3) | (uint32_t)(z & 7);
    return rd<float>(buf, leaf + 96 + (uint64_t)nf * 4);
}
inline float sample_nanovdb_density(const Medium& m, V3 p) {
    const HkMedium& h = m.h;
    float px = p.x - h.nanovdb_vec[0], py = p.y - h.nanovdb_vec[1], pz = p.z - h.nanovdb_vec[2];
    const float* M = h.nanovdb_inv_mat;
    float fxp = M[0] * px + M[1] * py + M[2] * pz;
    float fyp = M[3] * px + M[4] * py + M[5] * pz;
    float fzp = M[6] * px + M[7] * py + M[8] * pz;
    int32_t ix = floor_int32(fxp), iy = floor_int32(fyp), iz = floor_int32(fzp);
    float fx = fxp - (float)ix, fy = fyp - (float)iy, fz = fzp - (float)iz;
    float v000 = nanovdb_get_value(m, ix, iy, iz), v001 = nanovdb_get_value(m, ix, iy, iz + 1);
    float v010 = nanovdb_get_value(m, ix, iy + 1, iz), v011 = nanovdb_get_value(m, ix, iy + 1, iz + 1);
    float v100 = nanovdb_get_value(m, ix + 1, iy, iz), v101 = nanovdb_get_value(m, ix + 1, iy, iz + 1);
    float v110 = nanovdb_get_value(m, ix + 1, iy + 1, iz), v111 = nanovdb_get_value(m, ix + 1, iy + 1, iz + 1);
    float fx1 = 1.0f - fx, fy1 = 1.0f - fy, fz1 = 1.0f - fz;
    float v00 = v000 * fz1 + v001 * fz, v01 = v010 * fz1 + v011 * fz, v10 = v100 * fz1 + v101 * fz, v11 = v110 * fz1 + v111 * fz;
    float v0 = v00 * fy1 + v01 * fy, v1 = v10 * fy1 + v11 * fy;
    return v0 * fx1 + v1 * fx;
}
// media.jl:1544-1595
inline float sample_grid_density(const Medium& m, V3 pm) {
    const HkMedium& h = m.h;
    float pn[3];
    for (int k = 0; k < 3; k++) pn[k] = (pm[k] - h.bounds_min[k]) / (h.bounds_max[k] - h.bounds_min[k]);
    if (pn[0] < 0.0f || pn[1] < 0.0f || pn[2] < 0.0f || pn[0] > 1.0f || pn[1] > 1.0f || pn[2] > 1.0f) return 0.0f;
    int nx = h.density_res[0], ny = h.density_res[1], nz = h.density_res[2];
    float gx = pn[0] * (float)nx + 0.5f, gy = pn[1] * (float)ny + 0.5f, gz = pn[2] * (float)nz + 0.5f;
    int ix = clampi(floor_int32(gx), 1, nx - 1), iy = clampi(floor_int32(gy), 1, ny - 1), iz = clampi(floor_int32(gz), 1, nz - 1);
    float fx = clampf(gx - (float)ix, 0.0f, 1.0f), fy = clampf(gy - (float)iy, 0.0f, 1.0f), fz = clampf(gz - (float)iz, 0.0f, 1.0f);
    auto D = [&](int x, int y, int z) { return m.density[(size_t)(x - 1) + (size_t)nx * ((size_t)(y - 1) + (size_t)ny * (size_t)(z - 1))]; };
    float fx1 = 1.0f - fx;
    float d00 = D(ix, iy, iz) * fx1 + D(ix + 1, iy, iz) * fx;
    float d10 = D(ix, iy + 1, iz) * fx1 + D(ix + 1, iy + 1, iz) * fx;
    float d01 = D(ix, iy, iz + 1) * fx1 + D(ix + 1, iy, iz + 1) * fx;
    float d11 = D(ix, iy + 1, iz + 1) * fx1 + D(ix + 1, iy + 1, iz + 1) * fx;
    float fy1 = 1.0f - fy;
    float d0 = d00 * fy1 + d10 * fy, d1 = d01 * fy1 + d11 * fy;
    return d0 * (1.0f - fz) + d1 * fz;
}
// _sample_rgb_grid, media.jl:1283-1325 (RGB trilinear; zero outside [0,1]^3)
inline void sample_rgb_grid(const HkMedium& h, const std::vector<float>& grid, const float* pn, float* out) {
    out[0] = out[1] = out[2] = 0.0f;
    if (pn[0] < 0.0f || pn[1] < 0.0f || pn[2] < 0.0f || pn[0] > 1.0f || pn[1] > 1.0f || pn[2] > 1.0f) return;
    int nx = h.density_res[0], ny = h.density_res[1], nz = h.density_res[2];
    float gx = pn[0] * (float)nx + 0.5f, gy = pn[1] * (float)ny + 0.5f, gz = pn[2] * (float)nz + 0.5f;
    int ix = clampi(floor_int32(gx), 1, nx - 1), iy = clampi(floor_int32(gy), 1, ny - 1), iz = clampi(floor_int32(gz), 1, nz - 1);
    float fx = clampf(gx - (float)ix, 0.0f, 1.0f), fy = clampf(gy - (float)iy, 0.0f, 1.0f), fz = clampf(gz - (float)iz, 0.0f, 1.0f);
    auto G = [&](int x, int y, int z, int c) { return grid[3 * ((size_t)(x - 1) + (size_t)nx * ((size_t)(y - 1) + (size_t)ny * (size_t)(z - 1))) + c]; };
    float fx1 = 1.0f - fx, fy1 = 1.0f - fy;
    for (int c = 0; c < 3; c++) {
        float c00 = G(ix, iy, iz, c) * fx1 + G(ix + 1, iy, iz, c) * fx;
        float c10 = G(ix, iy + 1, iz, c) * fx1 + G(ix + 1, iy + 1, iz, c) * fx;
        float c01 = G(ix, iy, iz + 1, c) * fx1 + G(ix + 1, iy, iz + 1, c) * fx;
        float c11 = G(ix, iy + 1, iz + 1, c) * fx1 + G(ix + 1, iy + 1, iz + 1, c) * fx;
        float c0 = c00 * fy1 + c10 * fy, c1 = c01 * fy1 + c11 * fy;
        out[c] = c0 * (1.0f - fz) + c1 * fz;
    }
}
inline V3 affine_point(const float* M, V3 p) {   // rows 1-3 of a row-major 4x4, no divide (media.jl:1605-1608)
    return V3(M[0] * p.x + M[1] * p.y + M[2] * p.z + M[3], M[4] * p.x + M[5] * p.y + M[6] * p.z + M[7], M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11]);
}
inline V3 affine_vec(const float* M, V3 v) {
    return V3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z);
}
// sample_point: media.jl:781-793 (Homogeneous), :1597-1623 (Grid), nanovdb.jl:477-492 (NanoVDB)
inline MediumProps sample_point(const MediaCtx& C, uint32_t idx, V3 p, const Wavelengths& l) {
    const Medium& m = C.media[idx - 1];
    const HkMedium& h = m.h;
    MediumProps r;
    r.g = h.g;
    if (h.type == HK_MEDIUM_RGBGRID) {      // media.jl:1327-1372
        V3 pm = affine_point(h.medium_from_render, p);
        float pn[3];
        for (int k = 0; k < 3; k++) pn[k] = (pm[k] - h.bounds_min[k]) / (h.bounds_max[k] - h.bounds_min[k]);
        float a[3] = {1.0f, 1.0f, 1.0f}, b[3] = {1.0f, 1.0f, 1.0f};
        if (!m.rgb_a.empty()) sample_rgb_grid(h, m.rgb_a, pn, a);
        if (!m.rgb_s.empty()) sample_rgb_grid(h, m.rgb_s, pn, b);
        r.sigma_a = uplift_rgb_unbounded(*C.T, a, l) * h.scale;
        r.sigma_s = uplift_rgb_unbounded(*C.T, b, l) * h.scale;
        r.Le = Spec(0.0f);
        if (!m.rgb_le.empty() && h.Le_scale > 0.0f) { float e[3]; sample_rgb_grid(h, m.rgb_le, pn, e); r.Le = uplift_rgb_unbounded(*C.T, e, l) * h.Le_scale; }
        return r;
    }
    Spec sa = uplift_rgb_unbounded(*C.T, h.sigma_a_rgb, l), ss = uplift_rgb_unbounded(*C.T, h.sigma_s_rgb, l);
    if (h.type == HK_MEDIUM_HOMOGENEOUS) {
        r.sigma_a = sa; r.sigma_s = ss; r.Le = uplift_rgb_unbounded(*C.T, h.Le_rgb, l);
    } else if (h.type == HK_MEDIUM_GRID) {
        float d = sample_grid_density(m, affine_point(h.medium_from_render, p));
        r.sigma_a = sa * d; r.sigma_s = ss * d; r.Le = Spec(0.0f);
    } else {
        float d = sample_nanovdb_density(m, p);
        r.sigma_a = sa * d; r.sigma_s = ss * d; r.Le = Spec(0.0f);
    }
    return r;
}
// create_majorant_iterator: media.jl:811-850 (Homogeneous), :1625-1688 (Grid), nanovdb.jl:509-543 (NanoVDB)
inline MajIter create_majorant_iterator(const MediaCtx& C, uint32_t idx, V3 o, V3 d, float t_max, const Wavelengths& l) {
    const Medium& m = C.media[idx - 1];
    const HkMedium& h = m.h;
    Spec sa = uplift_rgb_unbounded(*C.T, h.sigma_a_rgb, l), ss = uplift_rgb_unbounded(*C.T, h.sigma_s_rgb, l);
    Spec st = h.type == HK_MEDIUM_RGBGRID ? Spec(1.0f) : sa + ss;      // RGBGrid: unit sigma_t, the scale is in the majorant grid (media.jl:1402)
    if (h.type == HK_MEDIUM_HOMOGENEOUS) {
        MajIter it{}; it.mode = (0.0f >= t_max) ? 0 : 1; it.sigma_t = st; it.t_min = 0.0f; it.t_max = t_max; it.hom_called = false;
        return it;
    }
    V3 ro = o, rd_ = d;
    if (h.type == HK_MEDIUM_GRID || h.type == HK_MEDIUM_RGBGRID) {
        ro = affine_point(h.medium_from_render, o); rd_ = affine_vec(h.medium_from_render, d);
        if (rd_.x * rd_.x + rd_.y * rd_.y + rd_.z * rd_.z < 1.0e-20f) return majiter_invalid();
    }
    float te, tx; ray_bounds_intersect(ro, rd_, h.bounds_min, h.bounds_max, te, tx);
    te = jmax(te, 0.0f); tx = jmin(tx, t_max);
    if (te >= tx) return majiter_invalid();
    return create_dda_iterator(m.majorant.data(), h.majorant_res, h.bounds_min, h.bounds_max, ro, rd_, te, tx, st);
}

// ---- LCG, delta-tracking.jl:28-58 ---------------------------------------------------------------------
inline uint64_t lcg_init(V3 o, V3 d, float t_max) {
    uint32_t ox, oy, oz, tm, dx, dy, dz;
    std::memcpy(&ox, &o.x, 4); std::memcpy(&oy, &o.y, 4); std::memcpy(&oz, &o.z, 4); std::memcpy(&tm, &t_max, 4);
    std::memcpy(&dx, &d.x, 4); std::memcpy(&dy, &d.y, 4); std::memcpy(&dz, &d.z, 4);
    uint64_t s1 = mix_bits((uint64_t)ox ^ ((uint64_t)oy << 16) ^ ((uint64_t)oz << 32) ^ (uint64_t)tm);
    uint64_t s2 = mix_bits((uint64_t)dx ^ ((uint64_t)dy << 16) ^ ((uint64_t)dz << 32));
    return s1 ^ s2;
}
inline float lcg_next(uint64_t& s) {
    s = s * 0x5DEECE66Dull + 11ull;
    float r = (float)(uint32_t)(s >> 32) * 2.3283064365386963e-10f;
    return std::min(r, ONE_MINUS_EPS);
}

struct DeltaResult {
    enum Event { ABSORBED, SCATTER, SURVIVED } event;
    Spec beta, r_u, r_l; V3 p; float g;
    Spec Le_add; bool has_Le;
};
// sample_medium_interaction! + sample_T_maj_loop! + sample_segment!, delta-tracking.jl:142-453
// (depth/max_depth gating of the survive/scatter outcomes is applied by the caller exactly as the reference does)
inline DeltaResult delta_track(const MediaCtx& C, uint32_t medium, V3 o, V3 d, float t_max, const Wavelengths& lam,
                               Spec beta, Spec r_u, Spec r_l, int32_t depth, int32_t max_depth, Spec* L_accum) {
    DeltaResult R; R.has_Le = false; R.g = 0.0f;
    uint64_t rng = lcg_init(o, d, t_max);
    MajIter it = create_majorant_iterator(C, medium, o, d, t_max, lam);
    V3 ray_d = d;
    for (int sgi = 0; sgi < 256; sgi++) {
        MajSeg seg;
        if (!ray_majorant_next(it, seg)) break;
        Spec smaj = seg.sigma_maj;
        float smaj0 = smaj.v[0];
        if (smaj0 < 1.0e-10f) continue;
        float t = seg.t_min, t_end = seg.t_max;
        V3 ray_o = o + ray_d * t;
        bool seg_done = false;
        for (int si = 0; si < 1024 && !seg_done; si++) {
            float u = lcg_next(rng);
            float dt = -dm_logf(std::max(1.0e-10f, 1.0f - u)) / smaj0;
            float ts = t + dt;
            if (ts >= t_end) {
                float dr = t_end - t;
                Spec Tm = exp(-dr * smaj);
                float T0 = Tm.v[0];
                if (T0 > 1.0e-10f) { beta = beta * Tm / T0; r_u = r_u * Tm / T0; r_l = r_l * Tm / T0; }
                seg_done = true; break;
            }
            Spec Tm = exp(-dt * smaj);
            V3 p = ray_o + ray_d * dt;
            MediumProps mp = sample_point(C, medium, p, lam);
            if (!is_black(mp.Le) && depth < max_depth) {
                float pr = smaj0 * Tm.v[0];
                if (pr > 1.0e-10f) {
                    Spec r_e = r_u * smaj * Tm / pr;
                    if (!is_black(r_e) && L_accum) *L_accum = *L_accum + beta * mp.sigma_a * Tm * mp.Le / (pr * average(r_e));
                }
            }
            float p_abs = mp.sigma_a.v[0] / smaj0, p_sc = mp.sigma_s.v[0] / smaj0;
            float ue = lcg_next(rng);
            if (ue < p_abs) { R.event = DeltaResult::ABSORBED; R.beta = Spec(0.0f); R.r_u = r_u; R.r_l = r_l; return R; }
            else if (ue < p_abs + p_sc) {
                if (depth >= max_depth) { R.event = DeltaResult::ABSORBED; R.beta = beta; R.r_u = r_u; R.r_l = r_l; return R; }
                float pdf = Tm.v[0] * mp.sigma_s.v[0];
                if (pdf > 1.0e-10f) { beta = beta * Tm * mp.sigma_s / pdf; r_u = r_u * Tm * mp.sigma_s / pdf; }
                R.event = DeltaResult::SCATTER; R.beta = beta; R.r_u = r_u; R.r_l = r_l; R.p = p; R.g = mp.g;
                return R;
            } else {
                Spec sn = smaj - mp.sigma_a - mp.sigma_s;
                sn = Spec(std::max(sn.v[0], 0.0f), std::max(sn.v[1], 0.0f), std::max(sn.v[2], 0.0f), std::max(sn.v[3], 0.0f));
                float pdf = Tm.v[0] * sn.v[0];
                if (pdf > 1.0e-10f) { beta = beta * Tm * sn / pdf; r_u = r_u * Tm * sn / pdf; r_l = r_l * Tm * smaj / pdf; }
                else { R.event = DeltaResult::ABSORBED; R.beta = Spec(0.0f); R.r_u = r_u; R.r_l = r_l; return R; }
                t = ts; ray_o = p;   // apply_deflection: identity (media.jl:2039)
                if (is_black(beta) || is_black(r_u)) { R.event = DeltaResult::ABSORBED; R.beta = beta; R.r_u = r_u; R.r_l = r_l; return R; }
            }
        }
    }
    R.event = DeltaResult::SURVIVED; R.beta = beta; R.r_u = r_u; R.r_l = r_l;
    return R;
}
// wrapper used by the render loop: applies delta-tracking.jl:198-201 gating
inline DeltaResult sample_medium_interaction(const MediaCtx& C, uint32_t medium, V3 o, V3 d, float t_max, const Wavelengths& lam,
                                             const Spec& beta, const Spec& r_u, const Spec& r_l, int32_t depth = 0, int32_t max_depth = 1 << 30, Spec* L_accum = nullptr) {
    DeltaResult R = delta_track(C, medium, o, d, t_max, lam, beta, r_u, r_l, depth, max_depth, L_accum);
    if (R.event == DeltaResult::SURVIVED && (is_black(R.beta) || is_black(R.r_u) || depth >= max_depth)) R.event = DeltaResult::ABSORBED;
    return R;
}

// ---- ratio tracking, intersection.jl:421-542 ----------------------------------------------------------
inline void transmittance_ratio_tracking(const MediaCtx& C, uint32_t medium, V3 o, V3 d, float t_max, const Wavelengths& lam, Spec& T_ray, Spec& r_u, Spec& r_l) {
    T_ray = Spec(1.0f); r_u = Spec(1.0f); r_l = Spec(1.0f);
    MajIter it = create_majorant_iterator(C, medium, o, d, t_max, lam);
    PCG32 rng = pcg32_init(pbrt_hash(o), pbrt_hash(d));
    for (int sgi = 0; sgi < 256; sgi++) {
        MajSeg seg;
        if (!ray_majorant_next(it, seg)) break;
        Spec smaj = seg.sigma_maj;
        float smaj0 = smaj.v[0];
        if (smaj0 < 1.0e-10f) continue;
        float t = seg.t_min, t_end = seg.t_max;
        for (int si = 0; si < 100; si++) {
            float u = pcg32_f32(rng);
            float dt = -dm_logf(std::max(1.0e-10f, 1.0f - u)) / smaj0;
            float ts = t + dt;
            if (ts >= t_end) {
                float dr = t_end - t;
                Spec Tm = exp(-dr * smaj);
                float T0 = Tm.v[0];
                if (T0 > 1.0e-10f) { T_ray = T_ray * Tm / T0; r_l = r_l * Tm / T0; r_u = r_u * Tm / T0; }
                break;
            }
            V3 p = o + d * ts;
            MediumProps mp = sample_point(C, medium, p, lam);
            Spec sn = smaj - mp.sigma_a - mp.sigma_s;
            sn = Spec(std::max(sn.v[0], 0.0f), std::max(sn.v[1], 0.0f), std::max(sn.v[2], 0.0f), std::max(sn.v[3], 0.0f));
            Spec Tm = exp(-dt * smaj);
            float pr = Tm.v[0] * smaj0;
            if (pr > 1.0e-10f) { T_ray = T_ray * Tm * sn / pr; r_l = r_l * Tm * smaj / pr; r_u = r_u * Tm * sn / pr; }
            else { T_ray = Spec(0.0f); return; }
            Spec Tr = T_ray / std::max(1.0e-10f, average(r_l + r_u));
            if (max_component(Tr) < 0.05f) {
                float q = 0.75f;
                float rr = pcg32_f32(rng);
                if (rr < q) { T_ray = Spec(0.0f); return; }
                T_ray = T_ray / (1.0f - q);
            }
            if (is_black(T_ray)) return;
            t = ts;
        }
        if (is_black(T_ray)) break;
    }
}

}  // namespace ok

