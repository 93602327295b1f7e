"""Host-side NanoVDB construction (stays on the host in the reference too):
build_nanovdb_from_dense (src/integrators/volpath/nanovdb.jl:602-858), build_nanovdb_majorant_grid (:1174-1235)
and the NanoVDBMedium(data; bounds, …) constructor (:940-1000).  The device only READS the buffer
(nanovdb_get_value, :315-388).  Byte offsets handed to the C ABI are 0-based.
"""
import numpy as np

from . import _abi as A

f32 = np.float32
LEAF_DIM, LOWER_DIM, UPPER_DIM = 8, 16, 32
LOWER_MASK, UPPER_MASK = 127, 4095
LEAF_SIZE, LOWER_NODE_SIZE, UPPER_NODE_SIZE = 2144, 1088 + 4096 * 8, 8256 + 32768 * 8
ROOT_HEADER, ROOT_TILE = 64, 32


def _root_key(c):
    x, y, z = (int(v) & 0xFFFFFFFF for v in c)
    return ((z >> 12) & 0x1FFFFF) | (((y >> 12) & 0x1FFFFF) << 21) | (((x >> 12) & 0x1FFFFF) << 42)


def _upper_off(c):
    x, y, z = (int(v) & 0xFFFFFFFF for v in c)
    return (((x >> 7) & 31) << 10) | (((y >> 7) & 31) << 5) | ((z >> 7) & 31)


def _lower_off(c):
    x, y, z = (int(v) & 0xFFFFFFFF for v in c)
    return (((x >> 3) & 15) << 8) | (((y >> 3) & 15) << 4) | ((z >> 3) & 15)


def build_nanovdb_from_dense(data_xyz, origin, extent, background=0.0):
    """data indexed [x, y, z].  Returns (buffer: np.uint8[...], metadata dict) with 0-based byte offsets."""
    data = np.asarray(data_xyz, dtype=f32)
    nx, ny, nz = data.shape
    dx, dy, dz = extent[0] / nx, extent[1] / ny, extent[2] / nz
    nb = [-(-n // LEAF_DIM) for n in (nx, ny, nz)]
    pad = np.full((nb[0] * 8, nb[1] * 8, nb[2] * 8), f32(background), dtype=f32)
    pad[:nx, :ny, :nz] = data
    blocks = pad.reshape(nb[0], 8, nb[1], 8, nb[2], 8).transpose(0, 2, 4, 1, 3, 5)   # [bx,by,bz,lx,ly,lz]
    active = (blocks != f32(background)).any(axis=(3, 4, 5))
    coords = np.argwhere(active)                                  # sorted lexicographically (x, y, z)
    leaf_coords = [tuple(int(v) * 8 for v in c) for c in coords]
    n_leaves = len(leaf_coords)
    if n_leaves == 0:
        raise ValueError("NanoVDB build: volume has no active voxels")
    lower_to_leaves, upper_to_lowers = {}, {}
    for li, c in enumerate(leaf_coords):
        lower_to_leaves.setdefault(tuple(v & ~LOWER_MASK for v in c), []).append(li)
    lower_bases = sorted(lower_to_leaves)
    for i, lb in enumerate(lower_bases):
        upper_to_lowers.setdefault(tuple(v & ~UPPER_MASK for v in lb), []).append(i)
    upper_bases = sorted(upper_to_lowers)
    n_low, n_up = len(lower_bases), len(upper_bases)
    root_size = ROOT_HEADER + n_up * ROOT_TILE
    up_sec, low_sec = n_up * UPPER_NODE_SIZE, n_low * LOWER_NODE_SIZE
    total = root_size + up_sec + low_sec + n_leaves * LEAF_SIZE
    buf = np.zeros(total, dtype=np.uint8)

    def w(off, val, dt):
        buf[off:off + np.dtype(dt).itemsize] = np.frombuffer(np.array([val], dtype=dt).tobytes(), dtype=np.uint8)

    def setbit(mask_off, n):
        buf[mask_off + (n >> 3)] |= np.uint8(1 << (n & 7))

    upper_pos = lambda i: root_size + i * UPPER_NODE_SIZE
    lower_pos = lambda i: root_size + up_sec + i * LOWER_NODE_SIZE
    leaf_pos = lambda i: root_size + up_sec + low_sec + i * LEAF_SIZE   # leaves already in sorted order
    for li, c in enumerate(leaf_coords):
        off = leaf_pos(li)
        bx, by, bz = (v // 8 for v in c)
        vals = blocks[bx, by, bz].reshape(512)                      # index = lx<<6 | ly<<3 | lz
        w(off, c[0], np.int32); w(off + 4, c[1], np.int32); w(off + 8, c[2], np.int32)
        buf[off + 12:off + 15] = 7
        bits = np.packbits((vals != f32(background)).astype(np.uint8), bitorder="little")
        buf[off + 16:off + 80] = bits
        w(off + 80, vals.min(), f32); w(off + 84, vals.max(), f32)
        buf[off + 96:off + 96 + 2048] = np.frombuffer(np.ascontiguousarray(vals).tobytes(), dtype=np.uint8)
    for i, lb in enumerate(lower_bases):
        off = lower_pos(i)
        for k in range(3):
            w(off + 4 * k, lb[k], np.int32); w(off + 12 + 4 * k, lb[k] + 127, np.int32)
        for li in lower_to_leaves[lb]:
            n = _lower_off(leaf_coords[li])
            setbit(off + 544, n); setbit(off + 32, n)
            w(off + 1088 + n * 8, leaf_pos(li) - off, np.int64)
    for i, ub in enumerate(upper_bases):
        off = upper_pos(i)
        for k in range(3):
            w(off + 4 * k, ub[k], np.int32); w(off + 12 + 4 * k, ub[k] + 4095, np.int32)
        for low_i in upper_to_lowers[ub]:
            n = _upper_off(lower_bases[low_i])
            setbit(off + 4128, n); setbit(off + 32, n)
            w(off + 8256 + n * 8, lower_pos(low_i) - off, np.int64)
    lc = np.asarray(leaf_coords)
    idx_min, idx_max = lc.min(0), lc.max(0) + 8
    for k in range(3):
        w(4 * k, int(idx_min[k]), np.int32); w(12 + 4 * k, int(idx_max[k]), np.int32)
    w(24, n_up, np.uint32); w(28, background, f32)
    for ti, ub in enumerate(upper_bases):
        t = ROOT_HEADER + ti * ROOT_TILE
        w(t, _root_key(ub), np.uint64); w(t + 8, upper_pos(ti), np.int64); w(t + 16, 1, np.uint32); w(t + 20, background, f32)
    meta = dict(
        world_min=tuple(f32(v) for v in origin), world_max=tuple(f32(origin[k] + extent[k]) for k in range(3)),
        inv_mat=(f32(1 / dx), f32(0), f32(0), f32(0), f32(1 / dy), f32(0), f32(0), f32(0), f32(1 / dz)),
        vec=(f32(origin[0] + dx / 2), f32(origin[1] + dy / 2), f32(origin[2] + dz / 2)),
        root_offset=0, upper_offset=upper_pos(0), lower_offset=lower_pos(0), leaf_offset=leaf_pos(0),
        leaf_count=n_leaves, lower_count=n_low, upper_count=n_up, root_table_size=n_up,
        index_min=tuple(int(v) for v in idx_min), index_max=tuple(int(v) for v in idx_max))
    return buf, meta


def build_nanovdb_majorant_grid(data_xyz, meta, bounds, res=(64, 64, 64)):
    """nanovdb.jl:1174-1220 evaluated on the dense source (tree values == dense values, background 0).
    Returns [rz][ry][rx]."""
    data = np.asarray(data_xyz, dtype=f32)
    lo, hi = np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32)
    diag = (hi - lo).astype(f32)
    ranges = []
    for k in range(3):
        r = res[k]
        i = np.arange(r, dtype=f32)
        pmin = (lo[k] + diag[k] * i / f32(r)).astype(f32)
        pmax = (lo[k] + diag[k] * (i + f32(1)) / f32(r)).astype(f32)
        imin = (f32(meta["inv_mat"][4 * k]) * (pmin - f32(meta["vec"][k]))).astype(f32)
        imax = (f32(meta["inv_mat"][4 * k]) * (pmax - f32(meta["vec"][k]))).astype(f32)
        n0 = np.maximum(np.floor(np.minimum(imin, imax) - f32(1)).astype(np.int64), meta["index_min"][k])
        n1 = np.minimum(np.ceil(np.maximum(imin, imax) + f32(1)).astype(np.int64), meta["index_max"][k])
        ranges.append((n0, n1))
    # voxels beyond the dense array (inside partially filled leaves) hold the background value 0
    cur = np.maximum(data, f32(0))
    for k in range(3):
        n0, n1 = ranges[k]
        n = cur.shape[k]
        out_shape = list(cur.shape)
        out_shape[k] = res[k]
        out = np.zeros(out_shape, dtype=f32)
        for i in range(res[k]):
            a, b = max(int(n0[i]), 0), min(int(n1[i]) + 1, n)
            if b > a:
                sl = [slice(None)] * 3
                sl[k] = slice(a, b)
                dst = [slice(None)] * 3
                dst[k] = i
                out[tuple(dst)] = cur[tuple(sl)].max(axis=k)
        cur = out
    return np.ascontiguousarray(cur.transpose(2, 1, 0))


class NanoVDBMedium:
    """NanoVDBMedium(data; bounds, σ_a, σ_s, g, majorant_res), nanovdb.jl:940-1000"""

    def __init__(self, data_xyz, bounds, sigma_a=0.0, sigma_s=1.0, g=0.0, majorant_res=(64, 64, 64)):
        from .host import _rgb
        self.bounds = (np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32))
        origin = [float(v) for v in self.bounds[0]]
        extent = [float(v) for v in (self.bounds[1] - self.bounds[0])]
        self.buffer, self.meta = build_nanovdb_from_dense(data_xyz, origin, extent)
        self.majorant_res = tuple(int(v) for v in majorant_res)
        self.majorant = build_nanovdb_majorant_grid(data_xyz, self.meta, self.bounds, self.majorant_res)
        self.sigma_a, self.sigma_s, self.g = _rgb(sigma_a), _rgb(sigma_s), float(g)

    def to_abi(self, keep):
        m = A.HkMedium(type=A.HK_MEDIUM_NANOVDB)
        m.sigma_a_rgb[:], m.sigma_s_rgb[:], m.Le_rgb[:] = self.sigma_a, self.sigma_s, (0, 0, 0)
        m.g, m.scale = self.g, 1.0
        m.bounds_min[:], m.bounds_max[:] = self.bounds[0].tolist(), self.bounds[1].tolist()
        ident = np.eye(4, dtype=f32).reshape(-1).tolist()
        m.render_from_medium[:], m.medium_from_render[:] = ident, ident
        m.majorant_res[:] = list(self.majorant_res)
        keep.append(self.majorant)
        m.majorant = self.majorant.ctypes.data_as(A.c_fp)
        keep.append(self.buffer)
        m.nanovdb_buf = self.buffer.ctypes.data_as(A.c_u8p)
        m.nanovdb_bytes = len(self.buffer)
        m.nanovdb_inv_mat[:] = [float(v) for v in self.meta["inv_mat"]]
        m.nanovdb_vec[:] = [float(v) for v in self.meta["vec"]]
        m.nanovdb_root_offset, m.nanovdb_upper_offset = self.meta["root_offset"], self.meta["upper_offset"]
        m.nanovdb_lower_offset, m.nanovdb_leaf_offset = self.meta["lower_offset"], self.meta["leaf_offset"]
        m.nanovdb_root_tiles, m.nanovdb_upper_count = self.meta["root_table_size"], self.meta["upper_count"]
        m.nanovdb_lower_count, m.nanovdb_leaf_count = self.meta["lower_count"], self.meta["leaf_count"]
        return m
