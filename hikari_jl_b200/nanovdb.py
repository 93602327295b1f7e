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


# ---- .nvdb files (nanovdb.jl:1058-1170) -----------------------------------------------------------------------------
# 0-based byte offsets of the fields the reference reads out of the NanoVDB GridData / TreeData headers (nanovdb.jl:9-32)
GRIDDATA_SIZE, TREEDATA_SIZE = 672, 64
MAP_INVMATF, MAP_VECF, WORLDBBOX = 296 + 36, 296 + 72, 560
TREE_NODE_OFFSETS, TREE_NODE_COUNTS = GRIDDATA_SIZE, GRIDDATA_SIZE + 32


def nanovdb_get_values(buf, meta, ijk):
    """nanovdb_get_value (nanovdb.jl:315-388) for an [N, 3] array of index coordinates, vectorised: root tile by key, upper and
    lower child masks / tables, leaf voxel; tile and background values where the tree has no child."""
    ijk = np.asarray(ijk, dtype=np.int64)
    xu, yu, zu = (ijk[:, k] & 0xFFFFFFFF for k in range(3))
    b = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    rd = lambda off, dt: np.ascontiguousarray(b[np.asarray(off, dtype=np.int64)[:, None] + np.arange(np.dtype(dt).itemsize)]).view(dt)[:, 0]
    root = int(meta["root_offset"])
    n = len(ijk)
    out = np.full(n, np.frombuffer(b[root + 28:root + 32].tobytes(), dtype=f32)[0], dtype=f32)       # background
    key = ((zu >> 12) & 0x1FFFFF) | (((yu >> 12) & 0x1FFFFF) << 21) | (((xu >> 12) & 0x1FFFFF) << 42)
    tile_off = np.full(n, -1, dtype=np.int64)
    for i in range(int(meta["root_table_size"])):                      # linear search, first match wins (:331-339)
        to = root + ROOT_HEADER + i * ROOT_TILE
        k = int(np.frombuffer(b[to:to + 8].tobytes(), dtype=np.uint64)[0])
        tile_off[(tile_off < 0) & (key == k)] = to
    live = np.flatnonzero(tile_off >= 0)
    if live.size == 0:
        return out
    child = rd(tile_off[live] + 8, np.int64)
    out[live[child == 0]] = rd(tile_off[live[child == 0]] + 20, f32) if (child == 0).any() else out[live[child == 0]]
    live, child = live[child != 0], child[child != 0]
    if live.size == 0:
        return out
    upper = root + child
    nu = (((xu[live] >> 7) & 31) << 10) | (((yu[live] >> 7) & 31) << 5) | ((zu[live] >> 7) & 31)
    on = ((b[upper + 4128 + (nu >> 3)] >> (nu & 7)) & 1) != 0
    if (~on).any():
        out[live[~on]] = rd(upper[~on] + 8256 + nu[~on] * 8, f32)
    live, upper, nu = live[on], upper[on], nu[on]
    if live.size == 0:
        return out
    lower = upper + rd(upper + 8256 + nu * 8, np.int64)
    nl = (((xu[live] >> 3) & 15) << 8) | (((yu[live] >> 3) & 15) << 4) | ((zu[live] >> 3) & 15)
    on = ((b[lower + 544 + (nl >> 3)] >> (nl & 7)) & 1) != 0
    if (~on).any():
        out[live[~on]] = rd(lower[~on] + 1088 + nl[~on] * 8, f32)
    live, lower, nl = live[on], lower[on], nl[on]
    if live.size == 0:
        return out
    leaf = lower + rd(lower + 1088 + nl * 8, np.int64)
    nf = ((ijk[live, 0] & 7) << 6) | ((ijk[live, 1] & 7) << 3) | (ijk[live, 2] & 7)
    out[live] = rd(leaf + 96 + nf * 4, f32)
    return out


def extract_nanovdb_metadata(buf):
    """nanovdb.jl:1109-1170; offsets returned 0-based."""
    b = np.asarray(buf, dtype=np.uint8)
    view = lambda off, dt, cnt: np.frombuffer(b[off:off + np.dtype(dt).itemsize * cnt].tobytes(), dtype=dt)
    bbox = view(WORLDBBOX, np.float64, 6)
    node_off, node_cnt = view(TREE_NODE_OFFSETS, np.uint64, 4), view(TREE_NODE_COUNTS, np.uint32, 3)
    leaf_off, lower_off, upper_off, root_off = (GRIDDATA_SIZE + int(v) for v in node_off)
    n_leaves = int(node_cnt[0])
    origins = np.stack([view(leaf_off + i * LEAF_SIZE, np.int32, 3) for i in range(n_leaves)]) if n_leaves else np.zeros((0, 3), np.int32)
    return dict(
        world_min=tuple(f32(v) for v in bbox[:3]), world_max=tuple(f32(v) for v in bbox[3:]),
        inv_mat=tuple(f32(v) for v in view(MAP_INVMATF, f32, 9)), vec=tuple(f32(v) for v in view(MAP_VECF, f32, 3)),
        root_offset=root_off, upper_offset=upper_off, lower_offset=lower_off, leaf_offset=leaf_off,
        leaf_count=n_leaves, lower_count=int(node_cnt[1]), upper_count=int(node_cnt[2]),
        root_table_size=int(view(root_off + 24, np.uint32, 1)[0]),
        index_min=tuple(int(v) for v in origins.min(axis=0)) if n_leaves else (2 ** 31 - 1,) * 3,
        index_max=tuple(int(v) + LEAF_DIM for v in origins.max(axis=0)) if n_leaves else (-2 ** 31,) * 3)


def parse_nanovdb_buffer(path):
    """parse_nanovdb_buffer(filepath) (nanovdb.jl:1085-1107): find the zlib stream in the first 500 bytes of the file, inflate it,
    read the metadata.  Returns (buffer np.uint8, metadata)."""
    import zlib
    raw = open(path, "rb").read()
    start = -1
    for i in range(min(500, len(raw) - 1)):
        if raw[i] == 0x78 and raw[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            start = i
            break
    if start < 0:
        raise ValueError("Could not find zlib header in NanoVDB file")
    buf = np.frombuffer(zlib.decompressobj().decompress(raw[start:]), dtype=np.uint8).copy()
    return buf, extract_nanovdb_metadata(buf)


def write_nanovdb_file(path, tree_buf, meta, header=b"NanoVDB0" + bytes(56)):
    """Test / export helper (no counterpart upstream): wrap a tree built by build_nanovdb_from_dense in GridData + TreeData headers
    laid out as extract_nanovdb_metadata expects, zlib-compress, and prefix a small uncompressed file header."""
    import zlib
    tree = np.asarray(tree_buf, dtype=np.uint8)
    # node sections as the builder lays them out: root | upper | lower | leaves, moved behind the two headers
    shift = GRIDDATA_SIZE + TREEDATA_SIZE
    out = np.zeros(shift + len(tree), dtype=np.uint8)
    out[shift:] = tree
    put = lambda off, arr: out.__setitem__(slice(off, off + arr.nbytes), np.frombuffer(arr.tobytes(), dtype=np.uint8))
    inv = np.asarray(meta["inv_mat"], dtype=f32)
    put(296, np.linalg.inv(inv.reshape(3, 3).astype(np.float64)).astype(f32).reshape(-1))
    put(MAP_INVMATF, inv); put(MAP_VECF, np.asarray(meta["vec"], dtype=f32))
    put(WORLDBBOX, np.asarray(list(meta["world_min"]) + list(meta["world_max"]), dtype=np.float64))
    rel = lambda k: np.uint64(TREEDATA_SIZE + int(meta[k]))          # offsets are relative to the TreeData start
    put(TREE_NODE_OFFSETS, np.array([rel("leaf_offset"), rel("lower_offset"), rel("upper_offset"), rel("root_offset")], dtype=np.uint64))
    put(TREE_NODE_COUNTS, np.array([meta["leaf_count"], meta["lower_count"], meta["upper_count"]], dtype=np.uint32))
    with open(path, "wb") as f:
        f.write(header)
        f.write(zlib.compress(out.tobytes(), 6))


def build_nanovdb_majorant_grid_from_buffer(buf, meta, bounds, res=(64, 64, 64)):
    """build_nanovdb_majorant_grid (nanovdb.jl:1174-1235) on a parsed buffer: per coarse cell, the maximum tree value over the index
    box spanned by the cell's two corners (+- 1 voxel), clipped to the leaves' index range.  Returns [rz][ry][rx]."""
    lo, hi = np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32)
    diag = (hi - lo).astype(f32)
    imin, imax = np.asarray(meta["index_min"], dtype=np.int64), np.asarray(meta["index_max"], dtype=np.int64)
    ext = np.maximum(imax - imin + 1, 0)
    dense = np.zeros(tuple(int(v) for v in ext), dtype=f32)
    if dense.size:                                     # every voxel of the clip range, through the same look-up the device uses
        gx, gy, gz = np.meshgrid(*(np.arange(imin[k], imax[k] + 1) for k in range(3)), indexing="ij")
        coords = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
        vals = np.concatenate([nanovdb_get_values(buf, meta, coords[i:i + (1 << 20)]) for i in range(0, len(coords), 1 << 20)])
        dense = vals.reshape(dense.shape)
    M = np.asarray(meta["inv_mat"], dtype=f32).reshape(3, 3)
    vec = np.asarray(meta["vec"], dtype=f32)
    def to_index(p):                                                       # world_to_index_f_raw (:1238-1246): (m1 px + m2 py) + m3 pz, f32
        q = (p - vec).astype(f32)
        return ((M[:, 0] * q[0] + M[:, 1] * q[1]).astype(f32) + M[:, 2] * q[2]).astype(f32)
    out = np.zeros((res[2], res[1], res[0]), dtype=f32)
    for iz in range(res[2]):
        for iy in range(res[1]):
            for ix in range(res[0]):
                c = np.array([ix, iy, iz], dtype=f32)
                p0 = (lo + diag * c / np.asarray(res, dtype=f32)).astype(f32)
                p1 = (lo + diag * (c + f32(1)) / np.asarray(res, dtype=f32)).astype(f32)
                i0, i1 = to_index(p0), to_index(p1)
                n0 = np.maximum(np.floor(np.minimum(i0, i1) - f32(1)).astype(np.int64), imin) - imin
                n1 = np.minimum(np.ceil(np.maximum(i0, i1) + f32(1)).astype(np.int64), imax) - imin
                if (n1 >= n0).all():
                    blk = dense[n0[0]:n1[0] + 1, n0[1]:n1[1] + 1, n0[2]:n1[2] + 1]
                    out[iz, iy, ix] = max(f32(0), blk.max()) if blk.size else f32(0)
    return out


class NanoVDBMedium:
    """NanoVDBMedium(data; bounds, σ_a, σ_s, g, majorant_res), nanovdb.jl:940-1000; NanoVDBMedium.from_file = the
    NanoVDBMedium(filepath; σ_a, σ_s, g, transform, majorant_res) constructor, :1320-1416"""

    @classmethod
    def from_file(cls, path, sigma_a=0.5, sigma_s=10.0, g=0.0, transform=None, majorant_res=(64, 64, 64)):
        from .host import _rgb
        self = cls.__new__(cls)
        self._built = None
        self.buffer, meta = parse_nanovdb_buffer(path)
        T = np.eye(3, dtype=f32) if transform is None else np.asarray(transform, dtype=f32).reshape(3, 3)
        inv_T = np.linalg.inv(T.astype(np.float64)).astype(f32)
        # world -> medium -> index (:1357-1358).  `vec` is NOT rotated, as in the reference (its comment :1350-1355: "vec is in medium
        # space ... for the bunny scene vec = (0,0,0), so this is fine")
        combined = (np.asarray(meta["inv_mat"], dtype=f32).reshape(3, 3) @ inv_T).astype(f32)
        wmin, wmax = np.asarray(meta["world_min"], dtype=f32), np.asarray(meta["world_max"], dtype=f32)
        corners = np.array([[(wmin, wmax)[(i >> 2) & 1][0], (wmin, wmax)[(i >> 1) & 1][1], (wmin, wmax)[i & 1][2]] for i in range(8)], dtype=f32)
        cw = (corners @ T.T).astype(f32)
        self.bounds = (cw.min(axis=0).astype(f32), cw.max(axis=0).astype(f32))
        self.meta = dict(meta, inv_mat=tuple(f32(v) for v in combined.reshape(-1)))
        self.majorant_res = tuple(int(v) for v in majorant_res)
        self._majorant, self._dense = None, None
        self.sigma_a, self.sigma_s, self.g = _rgb(sigma_a), _rgb(sigma_s), float(g)
        return self

    def __init__(self, data_xyz, bounds, sigma_a=0.0, sigma_s=1.0, g=0.0, majorant_res=(64, 64, 64)):
        from .host import _rgb
        self.bounds = (np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32))
        self.majorant_res = tuple(int(v) for v in majorant_res)
        self._majorant, self._dense = None, np.ascontiguousarray(np.asarray(data_xyz, dtype=f32))
        self._built = None                 # (buffer, meta) of the host builder, made on first use
        self.sigma_a, self.sigma_s, self.g = _rgb(sigma_a), _rgb(sigma_s), float(g)

    def _host_tree(self):
        """build_nanovdb_from_dense on the host (numpy): what the oracle is handed, what a .nvdb file is written from, and what the
        device-built tree is compared with byte for byte; the CUDA back end builds its own from the dense volume (nanovdb_buf = NULL)"""
        if self._built is None:
            origin = [float(v) for v in self.bounds[0]]
            extent = [float(v) for v in (self.bounds[1] - self.bounds[0])]
            self._built = build_nanovdb_from_dense(self._dense, origin, extent)
        return self._built

    @property
    def buffer(self):
        return self._host_tree()[0]

    @buffer.setter
    def buffer(self, b):                   # (from_file: the parsed file buffer)
        self._built = (b, (self._built or (None, None))[1])

    @property
    def meta(self):
        return self._host_tree()[1]

    @meta.setter
    def meta(self, m):
        self._built = ((self._built or (None, None))[0], m)

    def _index_transform(self):
        """inv_mat / vec of a tree built from the dense volume (build_nanovdb_from_dense's map: voxel centres at origin + (i + 1/2) d),
        without building the tree"""
        nx, ny, nz = self._dense.shape
        ext = [float(v) for v in (self.bounds[1] - self.bounds[0])]
        org = [float(v) for v in self.bounds[0]]
        dx, dy, dz = ext[0] / nx, ext[1] / ny, ext[2] / nz
        return ((f32(1 / dx), f32(0), f32(0), f32(0), f32(1 / dy), f32(0), f32(0), f32(0), f32(1 / dz)),
                (f32(org[0] + dx / 2), f32(org[1] + dy / 2), f32(org[2] + dz / 2)))

    @property
    def majorant(self):
        """build_nanovdb_majorant_grid on the host (numpy): what the oracle is handed, and what the device-built grid is tested
        against; the CUDA back end builds its own from the uploaded tree (HkMedium.majorant = NULL)"""
        if self._majorant is None:
            if self._dense is not None:
                self._majorant = build_nanovdb_majorant_grid(self._dense, self.meta, self.bounds, self.majorant_res)
            else:
                self._majorant = np.ascontiguousarray(build_nanovdb_majorant_grid_from_buffer(self.buffer, self.meta, self.bounds, self.majorant_res))
        return self._majorant

    def to_abi(self, keep, device_majorant=False):
        m = A.HkMedium(type=A.HK_MEDIUM_NANOVDB)
        m.sigma_a_rgb[:], m.sigma_s_rgb[:], m.Le_rgb[:] = self.sigma_a, self.sigma_s, (0, 0, 0)
        m.g, m.scale = self.g, 1.0
        m.bounds_min[:], m.bounds_max[:] = self.bounds[0].tolist(), self.bounds[1].tolist()
        ident = np.eye(4, dtype=f32).reshape(-1).tolist()
        m.render_from_medium[:], m.medium_from_render[:] = ident, ident
        m.majorant_res[:] = list(self.majorant_res)
        if not device_majorant:
            keep.append(self.majorant)
            m.majorant = self.majorant.ctypes.data_as(A.c_fp)
        if device_majorant and self._dense is not None:      # the CUDA back end: tree, majorant grid and dense mirror are all built on the device
            d = np.ascontiguousarray(self._dense.transpose(2, 1, 0))      # -> [nz][ny][nx]
            keep.append(d)
            m.density = d.ctypes.data_as(A.c_fp)
            m.density_res[:] = list(self._dense.shape)
            inv_mat, vec = self._index_transform()
            m.nanovdb_inv_mat[:] = [float(v) for v in inv_mat]
            m.nanovdb_vec[:] = [float(v) for v in vec]
            return m
        m.nanovdb_index_min[:] = [int(v) for v in self.meta["index_min"]]      # the clip range of the majorant build (nanovdb.jl:1137-1150)
        m.nanovdb_index_max[:] = [int(v) for v in self.meta["index_max"]]
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
