// hk_media.cuh — participating media on the device: HG phase, majorant DDA, Homogeneous / Grid / NanoVDB
// density, delta tracking and ratio tracking.
// Reference: src/integrators/volpath/media.jl:28-74,275-391,625-729,781-850,1544-1740, nanovdb.jl:246-543,
// delta-tracking.jl:142-453, intersection.jl:421-542.
#pragma once
#include "hk_spectral.cuh"

struct DevMedium {
    int32_t type; float sigma_a[3], sigma_s[3], Le[3]; float g;
    float bmin[3], bmax[3]; float medium_from_render[12];
    int32_t dres[3]; const float* __restrict__ density;
    int32_t mres[3]; const float* __restrict__ majorant;
    const uint8_t* __restrict__ nvdb; float inv_mat[9], vec[3]; uint64_t root_off; int32_t root_tiles;
    // RGBGridMedium (media.jl:1002-1456): per-voxel RGB sigma_a / sigma_s / Le, [nz][ny][nx][3], null = absent (sigma: 1, Le: 0)
    const float* __restrict__ rgb_a; const float* __restrict__ rgb_s; const float* __restrict__ rgb_le; float sigma_scale, le_scale;
    // one bit per majorant cell, set when the cell's majorant is exactly 0 (built on the device at upload, k_majorant_mask): such a
    // cell yields sigma_maj = sigma_t * 0 = 0 < 1e-10 for every wavelength, i.e. the tracking loops skip it -- the DDA can step over it
    // without fetching the grid value.  Bit i = cell x + rx (y + ry z).
    const uint32_t* __restrict__ maj_empty;
    // NanoVDB: a dense mirror of the tree's values over the index box [dn_min, dn_min + dn_ext) (built on the device at upload,
    // k_nvdb_densify, when the box fits HK_DENSE_MIRROR_MAX; z fastest like the leaves).  Look-ups whose eight corners fall inside the
    // box read the mirror -- the same values the root -> upper -> lower -> leaf walk returns -- everything else walks the tree.
    const float* __restrict__ dense; int32_t dn_min[3], dn_ext[3];
};
// smem_mask: the tracking kernels stage the empty-cell mask of ONE medium (smem_medium, 1-based; the first grid medium whose mask fits)
// in shared memory; other media read theirs from global memory
struct MediaCtx { DevTables T; const DevMedium* __restrict__ media; int32_t n_media; const uint32_t* smem_mask; int32_t smem_medium; };
#define HK_SMEM_MASK_WORDS 8192      // 64^3 cells = 32 KB

HK_DEV float hg_p(float g, float c) { float g2 = g * g, d = 1.0f + g2 - 2.0f * g * c; return (1.0f - g2) / (4.0f * HK_PI * d * sqrtf(d)); }   // media.jl:28-32
HK_DEV float3 sample_hg(float g, float3 wo, float2 u, float& pdf) {                                                                            // media.jl:42-74
    float c;
    if (fabsf(g) < 1.0e-3f) c = 1.0f - 2.0f * u.x;
    else { float g2 = g * g; float q = (1.0f - g2) / (1.0f - g + 2.0f * g * u.x); c = clampf((1.0f + g2 - q * q) / (2.0f * g), -1.0f, 1.0f); }
    float s = sqrtf(fmaxf(0.0f, 1.0f - c * c)), phi = 2.0f * HK_PI * u.y;
    Frame fr = make_frame(-wo);
    float3 wi = norm3(s * dm_cosf(phi) * fr.t + s * dm_sinf(phi) * fr.b + c * (-wo));
    pdf = hg_p(g, c);
    return wi;
}

// ---- majorant iterator ----------------------------------------------------------------------------------------
struct MajIter { int mode; Spec sigma_t; float t_min, t_max; bool hom_called; const float* __restrict__ grid; const uint32_t* mask; int res[3]; float next_t[3], delta_t[3]; int step[3], limit[3], voxel[3]; };
struct MajSeg { float t_min, t_max; Spec sigma_maj; };
HK_DEV float jl_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
HK_DEV float jl_min(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
HK_DEV void ray_bounds(float3 o, float3 d, const float* bmin, const float* bmax, float& te, float& tx) {   // media.jl:1704-1740
    float t0[3], t1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float dk = comp3(d, k), ok = comp3(o, k);
        float inv = fabsf(dk) > 1.0e-10f ? 1.0f / dk : (dk >= 0.0f ? HK_INF : -HK_INF);
        float a = (bmin[k] - ok) * inv, b = (bmax[k] - ok) * inv;
        if (a > b) { float t = a; a = b; b = t; }
        t0[k] = a; t1[k] = b;
    }
    te = jl_max(jl_max(t0[0], t0[1]), t0[2]);
    tx = jl_min(jl_min(t1[0], t1[1]), t1[2]);
}
HK_DEV void majiter_invalid(MajIter& it) { it.mode = 0; it.t_min = HK_INF; it.t_max = -HK_INF; it.hom_called = true; }
HK_DEV void dda_init(MajIter& it, const DevMedium& M, float3 o, float3 d, float t_min, float t_max, Spec sigma_t) {   // media.jl:275-391
    it.sigma_t = sigma_t; it.t_min = t_min; it.t_max = t_max; it.grid = M.majorant; it.hom_called = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int r = M.mres[k];
        it.res[k] = r;
        float diag = M.bmax[k] - M.bmin[k];
        float go = (comp3(o, k) - M.bmin[k]) / diag;
        float gd = comp3(d, k) * (fabsf(diag) > 1.0e-10f ? 1.0f / diag : 0.0f);
        float gi = go + gd * t_min;
        int vox = clampi(floor_i(gi * (float)r), 0, r - 1);
        it.delta_t[k] = fabsf(gd) > 1.0e-10f ? 1.0f / (fabsf(gd) * (float)r) : HK_INF;
        if (gd >= 0.0f) { it.next_t[k] = gd > 1.0e-10f ? t_min + ((float)(vox + 1) / (float)r - gi) / gd : HK_INF; it.step[k] = 1; it.limit[k] = r; }
        else { it.next_t[k] = gd < -1.0e-10f ? t_min + ((float)vox / (float)r - gi) / gd : HK_INF; it.step[k] = -1; it.limit[k] = -1; }
        it.voxel[k] = vox;
    }
    it.mode = t_min >= t_max ? 0 : 2;
}
// One DDA step over a cell whose majorant is known to be 0 (MajIter::mask): exactly the state changes of majiter_next() for that cell
// -- same axis choice, same t_min = min(next_t, t_max), same incremental next_t += delta_t, same end-of-grid handling -- without the grid
// fetch and without building the (unused) segment.  Returns 0: not applicable (no mask / cell not empty / iterator not in DDA mode):
// take majiter_next();  1: stepped over one empty cell;  2: the iterator is exhausted (majiter_next would have returned false).
HK_DEV int majiter_skip_empty(MajIter& it) {
    if (it.mode != 2 || it.mask == nullptr) return 0;
    const uint32_t cell = (uint32_t)(it.voxel[0] + it.res[0] * (it.voxel[1] + it.res[1] * it.voxel[2]));
    if (((it.mask[cell >> 5] >> (cell & 31u)) & 1u) == 0u) return 0;
    if (it.t_min >= it.t_max) { it.mode = 0; return 2; }
    const int axis = (it.next_t[0] < it.next_t[1]) ? ((it.next_t[0] < it.next_t[2]) ? 0 : 2) : ((it.next_t[1] < it.next_t[2]) ? 1 : 2);
    const float nt = axis == 0 ? it.next_t[0] : (axis == 1 ? it.next_t[1] : it.next_t[2]);
    it.t_min = fminf(nt, it.t_max);
    if (axis == 0) { it.voxel[0] += it.step[0]; it.next_t[0] += it.delta_t[0]; }
    else if (axis == 1) { it.voxel[1] += it.step[1]; it.next_t[1] += it.delta_t[1]; }
    else { it.voxel[2] += it.step[2]; it.next_t[2] += it.delta_t[2]; }
    if (it.voxel[0] == it.limit[0] || it.voxel[1] == it.limit[1] || it.voxel[2] == it.limit[2]) { it.mode = 0; it.t_min = it.t_max; }
    return 1;
}
HK_DEV bool majiter_next(MajIter& it, MajSeg& seg) {   // media.jl:625-729
    if (it.mode == 0) return false;
    if (it.mode == 1) {
        if (it.hom_called || it.t_min >= it.t_max) { it.mode = 0; return false; }
        seg.t_min = it.t_min; seg.t_max = it.t_max; seg.sigma_maj = it.sigma_t; it.hom_called = true;
        return true;
    }
    if (it.t_min >= it.t_max) { it.mode = 0; return false; }
    int axis = (it.next_t[0] < it.next_t[1]) ? ((it.next_t[0] < it.next_t[2]) ? 0 : 2) : ((it.next_t[1] < it.next_t[2]) ? 1 : 2);
    float nt = axis == 0 ? it.next_t[0] : (axis == 1 ? it.next_t[1] : it.next_t[2]);
    float st = fminf(nt, it.t_max);
    float rho = __ldg(it.grid + it.voxel[0] + it.res[0] * (it.voxel[1] + it.res[1] * it.voxel[2]));
    seg.t_min = it.t_min; seg.t_max = st; seg.sigma_maj = it.sigma_t * rho;
    it.t_min = st;
    if (axis == 0) { it.voxel[0] += it.step[0]; it.next_t[0] += it.delta_t[0]; }
    else if (axis == 1) { it.voxel[1] += it.step[1]; it.next_t[1] += it.delta_t[1]; }
    else { it.voxel[2] += it.step[2]; it.next_t[2] += it.delta_t[2]; }
    if (it.voxel[0] == it.limit[0] || it.voxel[1] == it.limit[1] || it.voxel[2] == it.limit[2]) { it.mode = 0; it.t_min = it.t_max; }
    return true;
}

// ---- NanoVDB (nanovdb.jl:246-469).  The 8 trilinear corners share one root->lower walk whenever they fall in
// the same 8^3 leaf (the common case); the per-thread LeafCache keeps that walk. Values read are exactly the
// ones the reference's 8 independent walks return. ------------------------------------------------------------
struct LeafCache { int32_t kx, ky, kz; uint64_t leaf_off; float tile; bool valid, is_leaf; };
template <class Tp> HK_DEV Tp rd(const uint8_t* __restrict__ b, uint64_t off) { return __ldg(reinterpret_cast<const Tp*>(b + off)); }
HK_DEV bool mask_on(const uint8_t* __restrict__ b, uint64_t off, uint32_t n) { return ((__ldg(b + off + (n >> 3)) >> (n & 7)) & 1) != 0; }
// root -> upper -> lower walk to the 8^3 leaf (or the tile value) that holds voxel (x, y, z): a real function -- nvdb_density has
// eight call sites and only the first of them (rarely the fifth) misses the per-thread leaf cache
struct LeafWalk { uint64_t leaf_off; float tile; bool is_leaf; };
HK_NI LeafWalk nvdb_walk(const uint8_t* __restrict__ b, uint64_t root_off, int32_t root_tiles, int32_t x, int32_t y, int32_t z) {
    LeafWalk r; r.leaf_off = 0; r.tile = 0.0f; r.is_leaf = false;
    uint32_t xu = (uint32_t)x, yu = (uint32_t)y, zu = (uint32_t)z;
    uint64_t key = (uint64_t)((zu >> 12) & 0x1fffff) | ((uint64_t)((yu >> 12) & 0x1fffff) << 21) | ((uint64_t)((xu >> 12) & 0x1fffff) << 42);
    uint64_t tile = 0; bool found = false;
    for (int i = 0; i < root_tiles; i++) { uint64_t to = root_off + 64 + (uint64_t)i * 32; if (rd<uint64_t>(b, to) == key) { found = true; tile = to; break; } }
    if (!found) { r.tile = rd<float>(b, root_off + 28); return r; }
    int64_t child = rd<int64_t>(b, tile + 8);
    if (child == 0) { r.tile = rd<float>(b, tile + 20); return r; }
    uint64_t upper = root_off + child;
    uint32_t nu = (((xu >> 7) & 31) << 10) | (((yu >> 7) & 31) << 5) | ((zu >> 7) & 31);
    if (!mask_on(b, upper + 4128, nu)) { r.tile = rd<float>(b, upper + 8256 + (uint64_t)nu * 8); return r; }
    uint64_t lower = upper + rd<int64_t>(b, upper + 8256 + (uint64_t)nu * 8);
    uint32_t nl = (((xu >> 3) & 15) << 8) | (((yu >> 3) & 15) << 4) | ((zu >> 3) & 15);
    if (!mask_on(b, lower + 544, nl)) { r.tile = rd<float>(b, lower + 1088 + (uint64_t)nl * 8); return r; }
    r.is_leaf = true; r.leaf_off = lower + rd<int64_t>(b, lower + 1088 + (uint64_t)nl * 8);
    return r;
}
HK_DEV float nvdb_value(const DevMedium& M, LeafCache& lc, int32_t x, int32_t y, int32_t z) {
    const int32_t kx = x >> 3, ky = y >> 3, kz = z >> 3;
    if (!(lc.valid && lc.kx == kx && lc.ky == ky && lc.kz == kz)) {
        const LeafWalk w = nvdb_walk(M.nvdb, M.root_off, M.root_tiles, x, y, z);
        lc.valid = true; lc.kx = kx; lc.ky = ky; lc.kz = kz; lc.is_leaf = w.is_leaf; lc.tile = w.tile; lc.leaf_off = w.leaf_off;
    }
    if (!lc.is_leaf) return lc.tile;
    uint32_t nf = ((uint32_t)(x & 7) << 6) | ((uint32_t)(y & 7) << 3) | (uint32_t)(z & 7);
    return rd<float>(M.nvdb, lc.leaf_off + 96 + (uint64_t)nf * 4);
}
HK_DEV void nvdb_index(const DevMedium& M, float3 p, float& gx, float& gy, float& gz) {      // index-from-world (nanovdb.jl:400-412)
    float px = p.x - M.vec[0], py = p.y - M.vec[1], pz = p.z - M.vec[2];
    gx = M.inv_mat[0] * px + M.inv_mat[1] * py + M.inv_mat[2] * pz;
    gy = M.inv_mat[3] * px + M.inv_mat[4] * py + M.inv_mat[5] * pz;
    gz = M.inv_mat[6] * px + M.inv_mat[7] * py + M.inv_mat[8] * pz;
}
HK_DEV float nvdb_lerp(float v000, float v001, float v010, float v011, float v100, float v101, float v110, float v111, float fx, float fy, float fz) {
    float fx1 = 1.0f - fx, fy1 = 1.0f - fy, fz1 = 1.0f - fz;
    float v00 = v000 * fz1 + v001 * fz, v01 = v010 * fz1 + v011 * fz, v10 = v100 * fz1 + v101 * fz, v11 = v110 * fz1 + v111 * fz;
    float v0 = v00 * fy1 + v01 * fy, v1 = v10 * fy1 + v11 * fy;
    return v0 * fx1 + v1 * fx;
}
// all eight trilinear corners of voxel (ix, iy, iz) inside the dense mirror?
HK_DEV bool nvdb_in_mirror(const DevMedium& M, int ix, int iy, int iz) {
    const uint32_t ux = (uint32_t)(ix - M.dn_min[0]), uy = (uint32_t)(iy - M.dn_min[1]), uz = (uint32_t)(iz - M.dn_min[2]);
    return M.dense != nullptr && ux < (uint32_t)(M.dn_ext[0] - 1) && uy < (uint32_t)(M.dn_ext[1] - 1) && uz < (uint32_t)(M.dn_ext[2] - 1);
}
HK_DEV float nvdb_mirror_lerp(const DevMedium& M, int ix, int iy, int iz, float fx, float fy, float fz) {
    const size_t sy = (size_t)M.dn_ext[2], sx = sy * (size_t)M.dn_ext[1];
    const float* __restrict__ b = M.dense + (size_t)(ix - M.dn_min[0]) * sx + (size_t)(iy - M.dn_min[1]) * sy + (size_t)(iz - M.dn_min[2]);
    return nvdb_lerp(__ldg(b), __ldg(b + 1), __ldg(b + sy), __ldg(b + sy + 1), __ldg(b + sx), __ldg(b + sx + 1), __ldg(b + sx + sy), __ldg(b + sx + sy + 1), fx, fy, fz);
}
// the eight corners through the tree (per-thread LeafCache: the corners share one root -> lower walk whenever they fall in one leaf)
HK_DEV float nvdb_tree_lerp(const DevMedium& M, LeafCache& lc, int ix, int iy, int iz, float fx, float fy, float fz) {
    float v000 = nvdb_value(M, lc, ix, iy, iz), v001 = nvdb_value(M, lc, ix, iy, iz + 1);
    float v010 = nvdb_value(M, lc, ix, iy + 1, iz), v011 = nvdb_value(M, lc, ix, iy + 1, iz + 1);
    float v100 = nvdb_value(M, lc, ix + 1, iy, iz), v101 = nvdb_value(M, lc, ix + 1, iy, iz + 1);
    float v110 = nvdb_value(M, lc, ix + 1, iy + 1, iz), v111 = nvdb_value(M, lc, ix + 1, iy + 1, iz + 1);
    return nvdb_lerp(v000, v001, v010, v011, v100, v101, v110, v111, fx, fy, fz);
}
// lc: the leaf that answered the previous look-up.  The trackers keep it across the events of a ray (consecutive events of a segment
// fall into the same 8^3 leaf almost always, so the root -> upper -> lower walk -- four dependent loads -- runs once per leaf, not once
// per event); a fresh cache (valid = false) gives the reference's literal behaviour.  Same values either way, and the same values from
// the dense mirror.
HK_DEV float nvdb_density(const DevMedium& M, float3 p, LeafCache& lc) {
    float gx, gy, gz;
    nvdb_index(M, p, gx, gy, gz);
    int ix = floor_i(gx), iy = floor_i(gy), iz = floor_i(gz);
    float fx = gx - (float)ix, fy = gy - (float)iy, fz = gz - (float)iz;
    if (nvdb_in_mirror(M, ix, iy, iz)) return nvdb_mirror_lerp(M, ix, iy, iz, fx, fy, fz);
    return nvdb_tree_lerp(M, lc, ix, iy, iz, fx, fy, fz);
}
HK_DEV float grid_density(const DevMedium& M, float3 pm) {   // media.jl:1544-1595
    float pn[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pn[k] = (comp3(pm, k) - M.bmin[k]) / (M.bmax[k] - M.bmin[k]);
    if (pn[0] < 0.0f || pn[1] < 0.0f || pn[2] < 0.0f || pn[0] > 1.0f || pn[1] > 1.0f || pn[2] > 1.0f) return 0.0f;
    const int nx = M.dres[0], ny = M.dres[1], nz = M.dres[2];
    float gx = pn[0] * (float)nx + 0.5f, gy = pn[1] * (float)ny + 0.5f, gz = pn[2] * (float)nz + 0.5f;
    int ix = clampi(floor_i(gx), 1, nx - 1), iy = clampi(floor_i(gy), 1, ny - 1), iz = clampi(floor_i(gz), 1, nz - 1);
    float fx = clampf(gx - (float)ix, 0.0f, 1.0f), fy = clampf(gy - (float)iy, 0.0f, 1.0f), fz = clampf(gz - (float)iz, 0.0f, 1.0f);
    const float* b = M.density + (size_t)(ix - 1) + (size_t)nx * ((size_t)(iy - 1) + (size_t)ny * (size_t)(iz - 1));
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    float fx1 = 1.0f - fx;
    float d00 = __ldg(b) * fx1 + __ldg(b + 1) * fx, d10 = __ldg(b + sy) * fx1 + __ldg(b + sy + 1) * fx;
    float d01 = __ldg(b + sz) * fx1 + __ldg(b + sz + 1) * fx, d11 = __ldg(b + sz + sy) * fx1 + __ldg(b + sz + sy + 1) * fx;
    float fy1 = 1.0f - fy;
    float d0 = d00 * fy1 + d10 * fy, d1 = d01 * fy1 + d11 * fy;
    return d0 * (1.0f - fz) + d1 * fz;
}
// _sample_rgb_grid, media.jl:1283-1325: trilinear RGB, zero outside the bounds
HK_DEV float3 rgbgrid_sample(const DevMedium& M, const float* __restrict__ grid, const float* pn) {
    if (pn[0] < 0.0f || pn[1] < 0.0f || pn[2] < 0.0f || pn[0] > 1.0f || pn[1] > 1.0f || pn[2] > 1.0f) return f3(0.0f, 0.0f, 0.0f);
    const int nx = M.dres[0], ny = M.dres[1], nz = M.dres[2];
    float gx = pn[0] * (float)nx + 0.5f, gy = pn[1] * (float)ny + 0.5f, gz = pn[2] * (float)nz + 0.5f;
    int ix = clampi(floor_i(gx), 1, nx - 1), iy = clampi(floor_i(gy), 1, ny - 1), iz = clampi(floor_i(gz), 1, nz - 1);
    float fx = clampf(gx - (float)ix, 0.0f, 1.0f), fy = clampf(gy - (float)iy, 0.0f, 1.0f), fz = clampf(gz - (float)iz, 0.0f, 1.0f);
    const float* b = grid + 3 * ((size_t)(ix - 1) + (size_t)nx * ((size_t)(iy - 1) + (size_t)ny * (size_t)(iz - 1)));
    const size_t sy = 3 * (size_t)nx, sz = 3 * (size_t)nx * ny;
    const float fx1 = 1.0f - fx, fy1 = 1.0f - fy;
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float c00 = __ldg(b + c) * fx1 + __ldg(b + 3 + c) * fx, c10 = __ldg(b + sy + c) * fx1 + __ldg(b + sy + 3 + c) * fx;
        float c01 = __ldg(b + sz + c) * fx1 + __ldg(b + sz + 3 + c) * fx, c11 = __ldg(b + sz + sy + c) * fx1 + __ldg(b + sz + sy + 3 + c) * fx;
        float c0 = c00 * fy1 + c10 * fy, c1 = c01 * fy1 + c11 * fy;
        out[c] = c0 * (1.0f - fz) + c1 * fz;
    }
    return f3(out[0], out[1], out[2]);
}
HK_DEV float3 affine_pt(const float* M, float3 p);
// sample_point(::RGBGridMedium), media.jl:1327-1372: two (three with emission) RGB look-ups and unbounded uplifts PER EVENT --
// kept out of line so the scalar-density media keep their register budget in the persistent tracking kernels
static __device__ __noinline__ void rgbgrid_props(const DevTables& T, const DevMedium& M, float3 p, float4 lam, Spec& sa, Spec& ss, Spec& Le) {
    const float3 pm = affine_pt(M.medium_from_render, p);
    float pn[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pn[k] = (comp3(pm, k) - M.bmin[k]) / (M.bmax[k] - M.bmin[k]);
    const float3 a = M.rgb_a ? rgbgrid_sample(M, M.rgb_a, pn) : f3(1.0f, 1.0f, 1.0f);
    const float3 b = M.rgb_s ? rgbgrid_sample(M, M.rgb_s, pn) : f3(1.0f, 1.0f, 1.0f);
    sa = uplift_rgb_unbounded(T, a.x, a.y, a.z, lam) * M.sigma_scale;
    ss = uplift_rgb_unbounded(T, b.x, b.y, b.z, lam) * M.sigma_scale;
    Le = sp(0.0f);
    if (M.rgb_le && M.le_scale > 0.0f) { const float3 e = rgbgrid_sample(M, M.rgb_le, pn); Le = uplift_rgb_unbounded(T, e.x, e.y, e.z, lam) * M.le_scale; }
}
HK_DEV float3 affine_pt(const float* M, float3 p) { return f3(M[0] * p.x + M[1] * p.y + M[2] * p.z + M[3], M[4] * p.x + M[5] * p.y + M[6] * p.z + M[7], M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11]); }
HK_DEV float3 affine_vc(const float* M, float3 v) { return f3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z); }

// per-ray cached coefficients: sigma_a / sigma_s spectra depend only on (medium, lambda)
struct MediumCoef { Spec sa, ss, Le; float g; };
HK_DEV MediumCoef medium_coef(const MediaCtx& C, const DevMedium& M, float4 lam) {
    MediumCoef c;
    if (C.T.med_pre) {      // uplift cache (DevTables)
        const float4* q = C.T.med_pre + 3 * (&M - (const DevMedium*)C.T.med_base);
        c.sa = pre_unbounded(__ldg(q), lam); c.ss = pre_unbounded(__ldg(q + 1), lam);
        c.Le = M.type == HK_MEDIUM_HOMOGENEOUS ? pre_unbounded(__ldg(q + 2), lam) : sp(0.0f);
    } else {
        c.sa = uplift_rgb_unbounded(C.T, M.sigma_a[0], M.sigma_a[1], M.sigma_a[2], lam);
        c.ss = uplift_rgb_unbounded(C.T, M.sigma_s[0], M.sigma_s[1], M.sigma_s[2], lam);
        c.Le = M.type == HK_MEDIUM_HOMOGENEOUS ? uplift_rgb_unbounded(C.T, M.Le[0], M.Le[1], M.Le[2], lam) : sp(0.0f);
    }
    c.g = M.g;
    return c;
}
HK_DEV float medium_density(const DevMedium& M, float3 p, LeafCache& lc) {
    if (M.type == HK_MEDIUM_GRID) return grid_density(M, affine_pt(M.medium_from_render, p));
    if (M.type == HK_MEDIUM_NANOVDB) return nvdb_density(M, p, lc);
    return 1.0f;
}
HK_DEV float medium_density(const DevMedium& M, float3 p) { LeafCache lc; lc.valid = false; return medium_density(M, p, lc); }
// The persistent tracking kernels keep each lane's LeafCache in SHARED memory between the events of its ray (lc_slot = the lane's
// column of HK_LC_WORDS words, stride blockDim.x; word 0 = 0 invalidates it): consecutive events of a segment fall into the same 8^3
// leaf almost always, so the root -> upper -> lower walk -- four dependent loads -- runs once per leaf instead of once per event,
// without the cache's seven words living in registers for the whole walk (registers decide this kernel's occupancy).
#define HK_LC_WORDS 7
// the tree path of medium_density_cached, out of line: with a dense mirror it only serves look-ups at the edge of / outside the tree's
// index box, and its eight cached look-ups are several hundred instructions the tracking loops should not have to fetch around
static __device__ __noinline__ float nvdb_density_tree_cached(const DevMedium& M, int ix, int iy, int iz, float fx, float fy, float fz, uint32_t* lc_slot) {
    const unsigned st = blockDim.x;
    LeafCache lc; lc.valid = false; lc.is_leaf = false; lc.kx = lc.ky = lc.kz = 0; lc.leaf_off = 0; lc.tile = 0.0f;
    if (lc_slot != nullptr) {
        const uint32_t fl = lc_slot[0];
        lc.valid = (fl & 1u) != 0u; lc.is_leaf = (fl & 2u) != 0u;
        lc.kx = (int32_t)lc_slot[st]; lc.ky = (int32_t)lc_slot[2 * st]; lc.kz = (int32_t)lc_slot[3 * st];
        lc.leaf_off = (uint64_t)lc_slot[4 * st] | ((uint64_t)lc_slot[5 * st] << 32); lc.tile = __uint_as_float(lc_slot[6 * st]);
    }
    const float v = nvdb_tree_lerp(M, lc, ix, iy, iz, fx, fy, fz);
    if (lc_slot != nullptr) {
        lc_slot[0] = (lc.valid ? 1u : 0u) | (lc.is_leaf ? 2u : 0u);
        lc_slot[st] = (uint32_t)lc.kx; lc_slot[2 * st] = (uint32_t)lc.ky; lc_slot[3 * st] = (uint32_t)lc.kz;
        lc_slot[4 * st] = (uint32_t)lc.leaf_off; lc_slot[5 * st] = (uint32_t)(lc.leaf_off >> 32); lc_slot[6 * st] = __float_as_uint(lc.tile);
    }
    return v;
}
HK_DEV float medium_density_cached(const DevMedium& M, float3 p, uint32_t* lc_slot) {
    if (M.type == HK_MEDIUM_GRID) return grid_density(M, affine_pt(M.medium_from_render, p));
    if (M.type != HK_MEDIUM_NANOVDB) return 1.0f;
    float gx, gy, gz;
    nvdb_index(M, p, gx, gy, gz);
    int ix = floor_i(gx), iy = floor_i(gy), iz = floor_i(gz);
    float fx = gx - (float)ix, fy = gy - (float)iy, fz = gz - (float)iz;
    if (nvdb_in_mirror(M, ix, iy, iz)) return nvdb_mirror_lerp(M, ix, iy, iz, fx, fy, fz);
    return nvdb_density_tree_cached(M, ix, iy, iz, fx, fy, fz, lc_slot);
}
HK_DEV void majiter_create(MajIter& it, const DevMedium& M, const MediumCoef& mc, float3 o, float3 d, float t_max, const uint32_t* mask = nullptr) {
    it.mask = mask;
    Spec st = M.type == HK_MEDIUM_RGBGRID ? sp(1.0f) : mc.sa + mc.ss;      // RGBGrid: the majorant grid already holds sigma_scale * max(sigma_a + sigma_s) (media.jl:1402)
    if (M.type == HK_MEDIUM_HOMOGENEOUS) { it.mode = (0.0f >= t_max) ? 0 : 1; it.sigma_t = st; it.t_min = 0.0f; it.t_max = t_max; it.hom_called = false; return; }
    float3 ro = o, rd_ = d;
    if (M.type == HK_MEDIUM_GRID || M.type == HK_MEDIUM_RGBGRID) {
        ro = affine_pt(M.medium_from_render, o); rd_ = affine_vc(M.medium_from_render, d);
        if (rd_.x * rd_.x + rd_.y * rd_.y + rd_.z * rd_.z < 1.0e-20f) { majiter_invalid(it); return; }
    }
    float te, tx; ray_bounds(ro, rd_, M.bmin, M.bmax, te, tx);
    te = jl_max(te, 0.0f); tx = jl_min(tx, t_max);
    if (te >= tx) { majiter_invalid(it); return; }
    dda_init(it, M, ro, rd_, te, tx, st);
}

// ---- delta tracking ----------------------------------------------------------------------------------------
#define HK_EV_ABSORBED 0
#define HK_EV_SCATTER 1
#define HK_EV_SURVIVED 2
// ray set-up of the trackers as a real function (to shrink the skip / event loop below the 32 KB instruction cache): tried and dropped --
// a member function that is not inlined takes `this` by address, which puts the whole tracker state (~100 words) in local memory
// (C4: 308 -> 262 Msamples/s)
#ifndef HK_NOINLINE_MEDIA_INIT
#define HK_NOINLINE_MEDIA_INIT 0
#endif
#if HK_NOINLINE_MEDIA_INIT
#define HK_MEDIA_INIT __device__ __noinline__
#else
#define HK_MEDIA_INIT HK_DEV
#endif
struct DeltaOut { int event; Spec beta, r_u, r_l; float3 p; float g; Spec Le_add; };
// Delta tracking as a state machine: step() performs ONE unit of work -- advance the majorant DDA to the next non-empty
// segment (skipping up to HK_TRACK_SKIP empty cells) or ONE tentative collision -- so a persistent warp can keep every lane
// on its own ray and hand finished lanes a new one (k_medium_track) instead of idling until the longest walk of the warp
// is done (ncu on C4: 7.9 of 32 lanes active in the one-ray-per-thread-to-completion form).  The arithmetic per ray, and
// therefore every output bit, is that of the nested loops in delta-tracking.jl:142-453.
// development statistics (-DHK_MEDIA_STATS): [0] rays set up, [1] empty cells stepped over, [2] cells fetched, [3] collision events,
// [4] segment ends, [5] skip-phase iterations, [6] event-phase iterations (per warp), [7] event-phase lane-iterations
#ifdef HK_MEDIA_STATS
__device__ unsigned long long g_media_stats[32];
#define HK_STAT(i, n) atomicAdd(&g_media_stats[i], (unsigned long long)(n))
#else
#define HK_STAT(i, n) ((void)0)
#endif
#ifndef HK_TRACK_SKIP
#ifndef HK_TRACK_SKIP
#define HK_TRACK_SKIP 4
#endif
#endif
#ifndef HK_TRACK_SKIP_EMPTY
#define HK_TRACK_SKIP_EMPTY 16
#endif
HK_DEV const uint32_t* medium_mask(const MediaCtx& C, int medium) {
    if (medium == C.smem_medium) return C.smem_mask;
    return C.media[medium - 1].maj_empty;
}
struct DeltaTracker {
    const DevMedium* M; MediumCoef mc; float3 o, d, ro; int depth, max_depth;
    uint64_t rng; MajIter it; Spec smaj; float seg_t_max, t; int sg, si; bool in_seg;
    Spec beta, r_u, r_l; DeltaOut R;
    HK_MEDIA_INIT void init(const MediaCtx& C, int medium, float3 o_, float3 d_, float t_max, float4 lam, Spec beta_, Spec r_u_, Spec r_l_, int depth_, int max_depth_) {
        R.g = 0.0f; R.p = f3(0, 0, 0); R.Le_add = sp(0.0f); R.event = HK_EV_SURVIVED; HK_STAT(0, 1);
        M = &C.media[medium - 1];
        mc = medium_coef(C, *M, lam);
        o = o_; d = d_; depth = depth_; max_depth = max_depth_; beta = beta_; r_u = r_u_; r_l = r_l_;
        rng = lcg_init(o, d, t_max);
        majiter_create(it, *M, mc, o, d, t_max, medium_mask(C, medium));
        sg = 0; si = 0; in_seg = false; t = 0.0f; ro = o; smaj = sp(0.0f); seg_t_max = 0.0f;
    }
    HK_DEV bool finish(int ev, Spec b) { R.event = ev; R.beta = b; R.r_u = r_u; R.r_l = r_l; return true; }
    // Both return true when the walk is over (R is final).  skip_step (precondition !in_seg): advance the DDA to the next
    // non-empty segment; event_step (precondition in_seg): one tentative collision.  The persistent kernels vote per warp
    // on which of the two to run, so the expensive event code executes with many lanes at once.
    // T / lam: only read for an RGBGridMedium (its coefficients are uplifted at every event); lam points at the path's wavelengths
    // RGB = false compiles the RGBGridMedium branch out: the persistent kernels are instantiated both ways and scenes without such
    // a medium run the lean one (with the branch in, k_medium_track needs 164 registers instead of 128)
    HK_DEV bool step(const DevTables& T, const float4* lam) { return in_seg ? event_step<true>(T, lam) : skip_step(); }
    HK_DEV bool skip_step() {
        for (int k = 0; k < HK_TRACK_SKIP && !in_seg; k++) {
            MajSeg seg;
            if (sg >= 256) return finish(HK_EV_SURVIVED, beta);
            {   // cells known to be empty cost a DDA step only (up to HK_TRACK_SKIP_EMPTY of them per unit of work)
                int r = 1;
                for (int e = 0; e < HK_TRACK_SKIP_EMPTY && sg < 256 && (r = majiter_skip_empty(it)) == 1; e++) { sg++; HK_STAT(1, 1); }
                if (r == 2 || sg >= 256) return finish(HK_EV_SURVIVED, beta);
                if (r == 1) continue;
            }
            if (!majiter_next(it, seg)) return finish(HK_EV_SURVIVED, beta);
            sg++; HK_STAT(2, 1);
            if (seg.sigma_maj.x < 1.0e-10f) continue;
            smaj = seg.sigma_maj; seg_t_max = seg.t_max; t = seg.t_min; ro = o + d * t; si = 0; in_seg = true;
        }
        return false;
    }
    template <bool RGB> HK_DEV bool event_step(const DevTables& T, const float4* lam, uint32_t* lc_slot = nullptr) {
        if (si >= 1024) { in_seg = false; return false; }
        si++;
        const float s0 = smaj.x;
        float u = lcg_next(rng);
        float dt = -dm_logf(fmaxf(1.0e-10f, 1.0f - u)) / s0;
        float ts = t + dt;
        if (ts >= seg_t_max) {
            HK_STAT(4, 1);
            Spec Tm = sp_exp(-(seg_t_max - t) * smaj);
            if (Tm.x > 1.0e-10f) { beta = beta * Tm / Tm.x; r_u = r_u * Tm / Tm.x; r_l = r_l * Tm / Tm.x; }
            in_seg = false;
            return false;
        }
        HK_STAT(3, 1);
        Spec Tm = sp_exp(-dt * smaj);
        float3 p = ro + d * dt;
        Spec sa, ss, Le = mc.Le;
        if (RGB && M->type == HK_MEDIUM_RGBGRID) rgbgrid_props(T, *M, p, *lam, sa, ss, Le);
        else {
            float dens = medium_density_cached(*M, p, lc_slot);
            sa = M->type == HK_MEDIUM_HOMOGENEOUS ? mc.sa : mc.sa * dens;
            ss = M->type == HK_MEDIUM_HOMOGENEOUS ? mc.ss : mc.ss * dens;
        }
        if (!sp_black(Le) && depth < max_depth) {
            float pr = s0 * Tm.x;
            if (pr > 1.0e-10f) {
                Spec re = r_u * smaj * Tm / pr;
                if (!sp_black(re)) R.Le_add = R.Le_add + beta * sa * Tm * Le / (pr * sp_avg(re));
            }
        }
        float pa = sa.x / s0, ps = ss.x / s0;
        float ue = lcg_next(rng);
        if (ue < pa) return finish(HK_EV_ABSORBED, sp(0.0f));
        if (ue < pa + ps) {
            if (depth >= max_depth) return finish(HK_EV_ABSORBED, beta);
            float pdf = Tm.x * ss.x;
            if (pdf > 1.0e-10f) { beta = beta * Tm * ss / pdf; r_u = r_u * Tm * ss / pdf; }
            R.p = p; R.g = mc.g;
            return finish(HK_EV_SCATTER, beta);
        }
        Spec sn = sp_max0(smaj - sa - ss);
        float pdf = Tm.x * sn.x;
        if (!(pdf > 1.0e-10f)) return finish(HK_EV_ABSORBED, sp(0.0f));
        beta = beta * Tm * sn / pdf; r_u = r_u * Tm * sn / pdf; r_l = r_l * Tm * smaj / pdf;
        t = ts; ro = p;
        if (sp_black(beta) || sp_black(r_u)) return finish(HK_EV_ABSORBED, beta);
        return false;
    }
};
// blocking form (stage-level parity tests, hk_test_delta_tracking)
HK_DEV DeltaOut delta_track(const MediaCtx& C, int medium, float3 o, float3 d, float t_max, float4 lam, Spec beta, Spec r_u, Spec r_l, int depth, int max_depth) {
    DeltaTracker T;
    T.init(C, medium, o, d, t_max, lam, beta, r_u, r_l, depth, max_depth);
    while (!T.step(C.T, &lam)) {}
    return T.R;
}

// ---- ratio tracking (shadow rays), intersection.jl:446-542, same state-machine form ---------------------------
struct RatioTracker {
    const DevMedium* M; MediumCoef mc; float3 o, d;
    Pcg32 rng; MajIter it; Spec smaj; float seg_t_max, t; int sg, si; bool in_seg;
    Spec T_ray, r_u, r_l;
    bool uni;      // grey medium: sigma_a / sigma_s are equal at the four wavelengths, so T_ray, r_u, r_l (which start at 1) stay equal in all
                   // four components; the spectral arithmetic of an event is then done once and broadcast -- the same operations on the same
                   // operands as each component would see, so the same bits, with 4 instead of 16 divisions per collision event
    HK_MEDIA_INIT void init(const MediaCtx& C, int medium, float3 o_, float3 d_, float t_max, float4 lam) {
        T_ray = sp(1.0f); r_u = sp(1.0f); r_l = sp(1.0f); HK_STAT(8, 1);
        M = &C.media[medium - 1];
        mc = medium_coef(C, *M, lam);
        uni = M->type != HK_MEDIUM_RGBGRID && mc.sa.x == mc.sa.y && mc.sa.y == mc.sa.z && mc.sa.z == mc.sa.w && mc.ss.x == mc.ss.y && mc.ss.y == mc.ss.z && mc.ss.z == mc.ss.w;
        o = o_; d = d_;
        majiter_create(it, *M, mc, o, d, t_max, medium_mask(C, medium));
        rng = pcg32_init(hash_f3(o), hash_f3(d));
        sg = 0; si = 0; in_seg = false; t = 0.0f; smaj = sp(0.0f); seg_t_max = 0.0f;
    }
    HK_DEV bool step(const DevTables& T, const float4* lam) { return in_seg ? event_step<true>(T, lam) : skip_step(); }
    HK_DEV bool skip_step() {
        for (int k = 0; k < HK_TRACK_SKIP && !in_seg; k++) {
            MajSeg seg;
            if (sg >= 256) return true;
            {
                int r = 1;
                for (int e = 0; e < HK_TRACK_SKIP_EMPTY && sg < 256 && (r = majiter_skip_empty(it)) == 1; e++) { sg++; HK_STAT(9, 1); }
                if (r == 2 || sg >= 256) return true;
                if (r == 1) continue;
            }
            if (!majiter_next(it, seg)) return true;
            sg++; HK_STAT(10, 1);
            if (seg.sigma_maj.x < 1.0e-10f) continue;
            smaj = seg.sigma_maj; seg_t_max = seg.t_max; t = seg.t_min; si = 0; in_seg = true;
        }
        return false;
    }
    template <bool RGB> HK_DEV bool event_step(const DevTables& T, const float4* lam, uint32_t* lc_slot = nullptr) {
        if (si >= 100) { in_seg = false; return sp_black(T_ray); }
        si++;
        const float s0 = smaj.x;
        float u = pcg32_f32(rng);
        float dt = -dm_logf(fmaxf(1.0e-10f, 1.0f - u)) / s0;
        float ts = t + dt;
        if (ts >= seg_t_max) {
            HK_STAT(12, 1);
            if (uni) {
                const float tm = dm_expf(-(seg_t_max - t) * smaj.x);
                if (tm > 1.0e-10f) { T_ray = sp(T_ray.x * tm / tm); r_l = sp(r_l.x * tm / tm); r_u = sp(r_u.x * tm / tm); }
                in_seg = false;
                return T_ray.x == 0.0f;
            }
            Spec Tm = sp_exp(-(seg_t_max - t) * smaj);
            if (Tm.x > 1.0e-10f) { T_ray = T_ray * Tm / Tm.x; r_l = r_l * Tm / Tm.x; r_u = r_u * Tm / Tm.x; }
            in_seg = false;
            return sp_black(T_ray);
        }
        HK_STAT(11, 1);
        float3 p = o + d * ts;
        Spec sa, ss;
        if (RGB && M->type == HK_MEDIUM_RGBGRID) { Spec le; rgbgrid_props(T, *M, p, *lam, sa, ss, le); }
        else {
            float dens = medium_density_cached(*M, p, lc_slot);
            sa = M->type == HK_MEDIUM_HOMOGENEOUS ? mc.sa : mc.sa * dens;
            ss = M->type == HK_MEDIUM_HOMOGENEOUS ? mc.ss : mc.ss * dens;
        }
        if (uni) {
            const float sn = fmaxf(smaj.x - sa.x - ss.x, 0.0f);
            const float tm = dm_expf(-dt * smaj.x);
            const float pr = tm * s0;
            if (!(pr > 1.0e-10f)) { T_ray = sp(0.0f); return true; }
            const float tr = T_ray.x * tm * sn / pr, rl = r_l.x * tm * smaj.x / pr, ru = r_u.x * tm * sn / pr;
            T_ray = sp(tr); r_l = sp(rl); r_u = sp(ru);
            const float sum = rl + ru;
            const float q = tr / fmaxf(1.0e-10f, (((sum + sum) + sum) + sum) / 4.0f);      // sp_maxc(T_ray / max(1e-10, sp_avg(r_l + r_u)))
            if (q < 0.05f) {
                if (pcg32_f32(rng) < 0.75f) { T_ray = sp(0.0f); return true; }
                T_ray = sp(tr / (1.0f - 0.75f));
            }
            if (T_ray.x == 0.0f) return true;
            t = ts;
            return false;
        }
        Spec sn = sp_max0(smaj - sa - ss);
        Spec Tm = sp_exp(-dt * smaj);
        float pr = Tm.x * s0;
        if (!(pr > 1.0e-10f)) { T_ray = sp(0.0f); return true; }
        T_ray = T_ray * Tm * sn / pr; r_l = r_l * Tm * smaj / pr; r_u = r_u * Tm * sn / pr;
        Spec Tr = T_ray / fmaxf(1.0e-10f, sp_avg(r_l + r_u));
        if (sp_maxc(Tr) < 0.05f) {
            if (pcg32_f32(rng) < 0.75f) { T_ray = sp(0.0f); return true; }
            T_ray = T_ray / (1.0f - 0.75f);
        }
        if (sp_black(T_ray)) return true;
        t = ts;
        return false;
    }
};
HK_DEV void ratio_track(const MediaCtx& C, int medium, float3 o, float3 d, float t_max, float4 lam, Spec& T_ray, Spec& r_u, Spec& r_l) {
    RatioTracker T;
    T.init(C, medium, o, d, t_max, lam);
    while (!T.step(C.T, &lam)) {}
    T_ray = T.T_ray; r_u = T.r_u; r_l = T.r_l;
}
