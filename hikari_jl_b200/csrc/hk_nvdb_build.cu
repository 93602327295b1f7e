// hk_nvdb_build.cu — build_nanovdb_from_dense (src/integrators/volpath/nanovdb.jl:602-858) on the device: a dense f32 volume
// becomes the NanoVDB buffer (root + upper 32^3 + lower 16^3 + leaf 8^3 nodes, the byte layout nanovdb_get_value :315-388 reads)
// without leaving HBM.  Active leaves = 8^3 blocks with any voxel != background, in lexicographic (x, y, z) block order -- the
// order the reference's sorted leaf list has -- so the buffer is byte for byte the one the host builder writes
// (hikari_jl_b200/nanovdb.py; tests/test_parity_gpu.py::test_nanovdb_tree_built_on_the_device).
//   k_nb_flags    one block per leaf candidate: active?  marks its lower / upper candidates, index bounding box (atomics)
//   cub scans     leaf / lower / upper candidate flags -> node numbers (lexicographic = the sorted order)
//   k_nb_leaves   one block per active leaf: header, value mask, min / max, 512 values; its bit + child offset in the lower node
//   k_nb_lowers   one thread per lower candidate: bounding box; its bit + child offset in the upper node
//   k_nb_uppers   one thread per upper candidate: bounding box, root tile; thread 0 writes the root header
#include "hk_context.h"
#include <cub/device/device_scan.cuh>
#include <climits>

namespace {
constexpr uint64_t NB_LEAF = 2144, NB_LOWER = 1088 + 4096 * 8, NB_UPPER = 8256 + 32768 * 8, NB_ROOT_HEADER = 64, NB_ROOT_TILE = 32;
struct NbDims { int n[3]; int nb[3]; int nl[3]; int nu[3]; };
struct NbOff { uint64_t upper, lower, leaf; };
__device__ inline uint32_t nb_lin(int x, int y, int z, const int* n) { return ((uint32_t)x * (uint32_t)n[1] + (uint32_t)y) * (uint32_t)n[2] + (uint32_t)z; }
__device__ inline float nb_voxel(const float* __restrict__ dens, const NbDims& D, int x, int y, int z, float bg) {      // dens is [nz][ny][nx]
    return (x < D.n[0] && y < D.n[1] && z < D.n[2]) ? __ldg(dens + ((size_t)z * D.n[1] + y) * D.n[0] + x) : bg;
}
__device__ inline void nb_setbit(uint8_t* mask, uint32_t n) { atomicOr(reinterpret_cast<uint32_t*>(mask) + (n >> 5), 1u << (n & 31u)); }

__global__ void __launch_bounds__(128) k_nb_flags(const float* __restrict__ dens, NbDims D, float bg, uint32_t* __restrict__ leaf_flag, uint32_t* __restrict__ low_flag,
                                                   uint32_t* __restrict__ up_flag, int32_t* __restrict__ bbox) {
    const uint32_t l = blockIdx.x;
    const int bz = (int)(l % (uint32_t)D.nb[2]), by = (int)((l / (uint32_t)D.nb[2]) % (uint32_t)D.nb[1]), bx = (int)(l / ((uint32_t)D.nb[2] * (uint32_t)D.nb[1]));
    int any = 0;
    for (int v = threadIdx.x; v < 512; v += blockDim.x) {      // x fastest: coalesced rows of 8
        const float val = nb_voxel(dens, D, bx * 8 + (v & 7), by * 8 + ((v >> 3) & 7), bz * 8 + (v >> 6), bg);
        any |= (val != bg) ? 1 : 0;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) {
        leaf_flag[l] = any ? 1u : 0u;
        if (any) {
            low_flag[nb_lin(bx >> 4, by >> 4, bz >> 4, D.nl)] = 1u;
            up_flag[nb_lin(bx >> 9, by >> 9, bz >> 9, D.nu)] = 1u;
            atomicMin(bbox + 0, bx * 8); atomicMin(bbox + 1, by * 8); atomicMin(bbox + 2, bz * 8);
            atomicMax(bbox + 3, bx * 8); atomicMax(bbox + 4, by * 8); atomicMax(bbox + 5, bz * 8);
        }
    }
}
__global__ void __launch_bounds__(128) k_nb_leaves(const float* __restrict__ dens, NbDims D, float bg, const uint32_t* __restrict__ leaf_flag, const uint32_t* __restrict__ leaf_idx,
                                                    const uint32_t* __restrict__ low_idx, uint8_t* __restrict__ buf, NbOff O) {
    __shared__ float vals[512];
    __shared__ float red[2][4];
    const uint32_t l = blockIdx.x;
    if (!leaf_flag[l]) return;
    const int bz = (int)(l % (uint32_t)D.nb[2]), by = (int)((l / (uint32_t)D.nb[2]) % (uint32_t)D.nb[1]), bx = (int)(l / ((uint32_t)D.nb[2] * (uint32_t)D.nb[1]));
    const uint64_t off = O.leaf + (uint64_t)leaf_idx[l] * NB_LEAF;
    float mn = 0.0f, mx = 0.0f; bool first = true;
    for (int v = threadIdx.x; v < 512; v += blockDim.x) {      // leaf voxel index = lx << 6 | ly << 3 | lz
        const float val = nb_voxel(dens, D, bx * 8 + (v >> 6), by * 8 + ((v >> 3) & 7), bz * 8 + (v & 7), bg);
        vals[v] = val;
        reinterpret_cast<float*>(buf + off + 96)[v] = val;
        mn = first ? val : fminf(mn, val); mx = first ? val : fmaxf(mx, val); first = false;
    }
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mn; red[1][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 16) {      // value mask: bit i <=> voxel i differs from the background
        uint32_t w = 0;
        for (int b = 0; b < 32; b++) w |= (vals[32 * threadIdx.x + b] != bg) ? (1u << b) : 0u;
        reinterpret_cast<uint32_t*>(buf + off + 16)[threadIdx.x] = w;
    }
    if (threadIdx.x == 0) {
        int32_t* h = reinterpret_cast<int32_t*>(buf + off);
        h[0] = bx * 8; h[1] = by * 8; h[2] = bz * 8;
        buf[off + 12] = 7; buf[off + 13] = 7; buf[off + 14] = 7;
        reinterpret_cast<float*>(buf + off + 80)[0] = fminf(fminf(red[0][0], red[0][1]), fminf(red[0][2], red[0][3]));
        reinterpret_cast<float*>(buf + off + 84)[0] = fmaxf(fmaxf(red[1][0], red[1][1]), fmaxf(red[1][2], red[1][3]));
        const uint64_t loff = O.lower + (uint64_t)low_idx[nb_lin(bx >> 4, by >> 4, bz >> 4, D.nl)] * NB_LOWER;
        const uint32_t n = ((uint32_t)(bx & 15) << 8) | ((uint32_t)(by & 15) << 4) | (uint32_t)(bz & 15);
        nb_setbit(buf + loff + 544, n); nb_setbit(buf + loff + 32, n);
        reinterpret_cast<int64_t*>(buf + loff + 1088)[n] = (int64_t)off - (int64_t)loff;
    }
}
__global__ void __launch_bounds__(128) k_nb_lowers(NbDims D, uint32_t n_cand, const uint32_t* __restrict__ low_flag, const uint32_t* __restrict__ low_idx,
                                                    const uint32_t* __restrict__ up_idx, uint8_t* __restrict__ buf, NbOff O) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand || !low_flag[i]) return;
    const int Lz = (int)(i % (uint32_t)D.nl[2]), Ly = (int)((i / (uint32_t)D.nl[2]) % (uint32_t)D.nl[1]), Lx = (int)(i / ((uint32_t)D.nl[2] * (uint32_t)D.nl[1]));
    const uint64_t off = O.lower + (uint64_t)low_idx[i] * NB_LOWER;
    int32_t* h = reinterpret_cast<int32_t*>(buf + off);
    h[0] = Lx * 128; h[1] = Ly * 128; h[2] = Lz * 128; h[3] = Lx * 128 + 127; h[4] = Ly * 128 + 127; h[5] = Lz * 128 + 127;
    const uint64_t uoff = O.upper + (uint64_t)up_idx[nb_lin(Lx >> 5, Ly >> 5, Lz >> 5, D.nu)] * NB_UPPER;
    const uint32_t n = ((uint32_t)(Lx & 31) << 10) | ((uint32_t)(Ly & 31) << 5) | (uint32_t)(Lz & 31);
    nb_setbit(buf + uoff + 4128, n); nb_setbit(buf + uoff + 32, n);
    reinterpret_cast<int64_t*>(buf + uoff + 8256)[n] = (int64_t)off - (int64_t)uoff;
}
__global__ void __launch_bounds__(128) k_nb_uppers(NbDims D, uint32_t n_cand, const uint32_t* __restrict__ up_flag, const uint32_t* __restrict__ up_idx, uint8_t* __restrict__ buf,
                                                    NbOff O, const int32_t* __restrict__ bbox, uint32_t n_up, float bg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {      // RootData header: index bounding box of the leaves, table size, background
        int32_t* r = reinterpret_cast<int32_t*>(buf);
        r[0] = bbox[0]; r[1] = bbox[1]; r[2] = bbox[2]; r[3] = bbox[3] + 8; r[4] = bbox[4] + 8; r[5] = bbox[5] + 8;
        reinterpret_cast<uint32_t*>(buf)[6] = n_up; reinterpret_cast<float*>(buf)[7] = bg;
    }
    if (i >= n_cand || !up_flag[i]) return;
    const int Uz = (int)(i % (uint32_t)D.nu[2]), Uy = (int)((i / (uint32_t)D.nu[2]) % (uint32_t)D.nu[1]), Ux = (int)(i / ((uint32_t)D.nu[2] * (uint32_t)D.nu[1]));
    const uint32_t ui = up_idx[i];
    const uint64_t off = O.upper + (uint64_t)ui * NB_UPPER;
    int32_t* h = reinterpret_cast<int32_t*>(buf + off);
    h[0] = Ux * 4096; h[1] = Uy * 4096; h[2] = Uz * 4096; h[3] = Ux * 4096 + 4095; h[4] = Uy * 4096 + 4095; h[5] = Uz * 4096 + 4095;
    uint8_t* t = buf + NB_ROOT_HEADER + (uint64_t)ui * NB_ROOT_TILE;      // root tile: key, child offset (from the root = buffer start), state, value
    reinterpret_cast<uint64_t*>(t)[0] = ((uint64_t)Uz & 0x1FFFFFull) | (((uint64_t)Uy & 0x1FFFFFull) << 21) | (((uint64_t)Ux & 0x1FFFFFull) << 42);
    reinterpret_cast<int64_t*>(t)[1] = (int64_t)off;
    reinterpret_cast<uint32_t*>(t)[4] = 1u; reinterpret_cast<float*>(t)[5] = bg;
}
}  // namespace

#define NB_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return HK_ERR_CUDA; } } while (0)

// dens_dev: [nz][ny][nx] f32 on the device.  On success `tree` holds the NanoVDB buffer and `out` its offsets / counts.
int32_t hk_nvdb_build_dense(HkContext* ctx, const float* dens_dev, const int32_t res[3], float background, DevBuf& tree, HkNvdbBuilt& out) {
    NbDims D;
    size_t nbt = 1, nlt = 1, nut = 1;
    for (int k = 0; k < 3; k++) {
        if (res[k] < 1 || res[k] > (1 << 20)) { ctx->err = "NanoVDB build: bad dense resolution"; return HK_ERR_INVALID; }
        D.n[k] = res[k]; D.nb[k] = (res[k] + 7) / 8; D.nl[k] = (D.nb[k] + 15) / 16; D.nu[k] = (D.nl[k] + 31) / 32;
        nbt *= (size_t)D.nb[k]; nlt *= (size_t)D.nl[k]; nut *= (size_t)D.nu[k];
    }
    if (nbt >= (1ull << 31)) { ctx->err = "NanoVDB build: volume too large"; return HK_ERR_INVALID; }
    cudaStream_t st = ctx->stream;
    // scratch: flags + scanned indices of the three candidate levels, the bounding box, cub's temporary storage
    const size_t n_all = nbt + nlt + nut;
    DevBuf scratch, tmp;
    NB_CK(scratch.alloc(4 * (2 * n_all + 8)));
    uint32_t* leaf_flag = scratch.as<uint32_t>(); uint32_t* low_flag = leaf_flag + nbt; uint32_t* up_flag = low_flag + nlt;
    uint32_t* leaf_idx = up_flag + nut; uint32_t* low_idx = leaf_idx + nbt; uint32_t* up_idx = low_idx + nlt;
    int32_t* bbox = reinterpret_cast<int32_t*>(up_idx + nut);
    NB_CK(cudaMemsetAsync(scratch.p, 0, scratch.bytes, st));
    const int32_t bbox0[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    NB_CK(cudaMemcpyAsync(bbox, bbox0, sizeof(bbox0), cudaMemcpyHostToDevice, st));
    k_nb_flags<<<(unsigned)nbt, 128, 0, st>>>(dens_dev, D, background, leaf_flag, low_flag, up_flag, bbox);
    size_t tmp_bytes = 0, need = 0;
    const size_t counts[3] = {nbt, nlt, nut};
    for (int k = 0; k < 3; k++) { cub::DeviceScan::ExclusiveSum(nullptr, need, leaf_flag, leaf_idx, (int)counts[k], st); tmp_bytes = std::max(tmp_bytes, need); }
    NB_CK(tmp.alloc(tmp_bytes));
    uint32_t* flags[3] = {leaf_flag, low_flag, up_flag}; uint32_t* idx[3] = {leaf_idx, low_idx, up_idx};
    for (int k = 0; k < 3; k++) { size_t tb = tmp.bytes; NB_CK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags[k], idx[k], (int)counts[k], st)); }
    uint32_t n_nodes[3];
    for (int k = 0; k < 3; k++) {
        uint32_t last_idx = 0, last_flag = 0;
        NB_CK(cudaMemcpyAsync(&last_idx, idx[k] + counts[k] - 1, 4, cudaMemcpyDeviceToHost, st));
        NB_CK(cudaMemcpyAsync(&last_flag, flags[k] + counts[k] - 1, 4, cudaMemcpyDeviceToHost, st));
        NB_CK(cudaStreamSynchronize(st));
        n_nodes[k] = last_idx + last_flag;
    }
    const uint32_t n_leaf = n_nodes[0], n_low = n_nodes[1], n_up = n_nodes[2];
    if (n_leaf == 0) { ctx->err = "NanoVDB build: volume has no active voxels"; scratch.release(); tmp.release(); return HK_ERR_INVALID; }
    const uint64_t root_size = NB_ROOT_HEADER + (uint64_t)n_up * NB_ROOT_TILE;
    NbOff O; O.upper = root_size; O.lower = O.upper + (uint64_t)n_up * NB_UPPER; O.leaf = O.lower + (uint64_t)n_low * NB_LOWER;
    const uint64_t total = O.leaf + (uint64_t)n_leaf * NB_LEAF;
    NB_CK(tree.alloc((size_t)total));
    NB_CK(cudaMemsetAsync(tree.p, 0, (size_t)total, st));
    uint8_t* buf = tree.as<uint8_t>();
    k_nb_leaves<<<(unsigned)nbt, 128, 0, st>>>(dens_dev, D, background, leaf_flag, leaf_idx, low_idx, buf, O);
    k_nb_lowers<<<(unsigned)((nlt + 127) / 128), 128, 0, st>>>(D, (uint32_t)nlt, low_flag, low_idx, up_idx, buf, O);
    k_nb_uppers<<<(unsigned)((nut + 127) / 128), 128, 0, st>>>(D, (uint32_t)nut, up_flag, up_idx, buf, O, bbox, n_up, background);
    ctx->launches += 7;
    int32_t bb[6];
    NB_CK(cudaMemcpyAsync(bb, bbox, sizeof(bb), cudaMemcpyDeviceToHost, st));
    NB_CK(cudaStreamSynchronize(st));
    NB_CK(cudaGetLastError());
    out.bytes = total; out.root_off = 0; out.upper_off = O.upper; out.lower_off = O.lower; out.leaf_off = O.leaf;
    out.n_up = (int32_t)n_up; out.n_low = (int32_t)n_low; out.n_leaf = (int32_t)n_leaf;
    for (int k = 0; k < 3; k++) { out.idx_min[k] = bb[k]; out.idx_max[k] = bb[3 + k] + 8; }
    scratch.release(); tmp.release();
    return HK_OK;
}
