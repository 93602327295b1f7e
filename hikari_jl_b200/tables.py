"""Constant tables of the VolPath path (host side).

* data/hikari_tables.npz — Sobol matrices, CIE 1931, D65, metal spectra, extracted by tools/extract_tables.py.
* the sRGB sigmoid-polynomial table (src/spectral/srgb_spectrum_table.dat upstream, MISSING from the checkout):
  regenerated once with hk_host_generate_rgb2spec (csrc/host_rgb2spec.cpp, a port of rgb2spec_gen.jl:332-409)
  and cached in data/srgb_spectrum_table.dat in the reference's own file format (rgb2spec.jl:403-412).
"""
import ctypes as C
import hashlib
import os
import numpy as np

from . import _abi as A

f32 = np.float32
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_NPZ = os.path.join(_ROOT, "data", "hikari_tables.npz")
_DAT = os.path.join(_ROOT, "data", "srgb_spectrum_table.dat")
_tables = None
_srgb = None


def load_tables():
    global _tables
    if _tables is None:
        z = np.load(_NPZ)
        _tables = {k: z[k] for k in z.files}
    return _tables


def generate_srgb_table(res=64):
    t = load_tables()
    lib = A.load_library()
    d65 = np.ascontiguousarray(t["gen_d65_raw"] / t["gen_d65_norm"][0])
    cx, cy, cz = (np.ascontiguousarray(t[k]) for k in ("gen_cie_x", "gen_cie_y", "gen_cie_z"))
    scale = np.zeros(res, dtype=f32)
    coeffs = np.zeros(9 * res ** 3, dtype=f32)
    dp = C.POINTER(C.c_double)
    rc = lib.hk_host_generate_rgb2spec(res, cx.ctypes.data_as(dp), cy.ctypes.data_as(dp), cz.ctypes.data_as(dp),
                                       d65.ctypes.data_as(dp), scale.ctypes.data_as(A.c_fp), coeffs.ctypes.data_as(A.c_fp))
    if rc != 0:
        raise RuntimeError("hk_host_generate_rgb2spec failed")
    return scale, coeffs


def get_srgb_table():
    """get_srgb_table(), rgb2spec.jl:424-441: load the cached .dat, else generate + save."""
    global _srgb
    if _srgb is None:
        if not os.path.exists(_DAT):
            scale, coeffs = generate_srgb_table(64)
            tmp = _DAT + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(np.int32(64).tobytes())
                f.write(scale.tobytes())
                f.write(coeffs.tobytes())
            os.replace(tmp, _DAT)
        raw = np.fromfile(_DAT, dtype=np.uint8)
        res = int(np.frombuffer(raw[:4].tobytes(), dtype=np.int32)[0])
        scale = np.frombuffer(raw[4:4 + 4 * res].tobytes(), dtype=f32).copy()
        coeffs = np.frombuffer(raw[4 + 4 * res:].tobytes(), dtype=f32).copy()
        assert coeffs.size == 9 * res ** 3
        _srgb = (scale, coeffs)
    return _srgb


def srgb_table_sha256():
    get_srgb_table()
    with open(_DAT, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _coef(coeffs, res, maxc, zi, yi, xi, c):
    return coeffs[(maxc - 1) + 3 * ((zi - 1) + res * ((yi - 1) + res * ((xi - 1) + res * (c - 1))))]


def rgb_to_spectrum(r, g, b):
    """Host-side rgb_to_spectrum (rgb2spec.jl:83-167) in Float32 — used only by light constructors
    (rgb_illuminant_spectrum, rgb2spec.jl:371-384) when a scene is built."""
    scale, coeffs = get_srgb_table()
    res = len(scale)
    r, g, b = (f32(min(max(v, 0.0), 1.0)) for v in (r, g, b))
    if r == g and g == b:
        if 0 < r < 1:
            c2 = f32(f32(r - f32(0.5)) / f32(np.sqrt(f32(r * f32(f32(1) - r)))))
        elif r <= 0:
            c2 = f32(-1e10)
        else:
            c2 = f32(1e10)
        return (f32(0), f32(0), c2)
    maxc = (1 if r > b else 3) if r > g else (2 if g > b else 3)
    z = r if maxc == 1 else (g if maxc == 2 else b)
    xc = g if maxc == 1 else (b if maxc == 2 else r)
    yc = b if maxc == 1 else (r if maxc == 2 else g)
    x = f32(f32(xc * f32(res - 1)) / z)
    y = f32(f32(yc * f32(res - 1)) / z)
    zi = 1
    for i in range(1, res):
        if scale[i - 1] < z:
            zi = i
    zi = min(zi, res - 1)
    xi = min(int(x) + 1, res - 1)
    yi = min(int(y) + 1, res - 1)
    dx = f32(x - f32(xi - 1))
    dy = f32(y - f32(yi - 1))
    dz = f32(f32(z - scale[zi - 1]) / f32(scale[zi] - scale[zi - 1]))
    one = f32(1)
    out = []
    for k in (1, 2, 3):
        cf = lambda a, b_, c_: _coef(coeffs, res, maxc, a, b_, c_, k)
        lo = f32(f32(one - dy) * f32(f32(one - dx) * cf(zi, yi, xi) + dx * cf(zi, yi, xi + 1))
                 + dy * f32(f32(one - dx) * cf(zi, yi + 1, xi) + dx * cf(zi, yi + 1, xi + 1)))
        hi = f32(f32(one - dy) * f32(f32(one - dx) * cf(zi + 1, yi, xi) + dx * cf(zi + 1, yi, xi + 1))
                 + dy * f32(f32(one - dx) * cf(zi + 1, yi + 1, xi) + dx * cf(zi + 1, yi + 1, xi + 1)))
        out.append(f32(f32(one - dz) * lo + dz * hi))
    return tuple(out)


def rgb_illuminant_spectrum(rgb):
    """rgb_illuminant_spectrum(table, r, g, b) -> (poly, scale), rgb2spec.jl:371-384"""
    r, g, b = (f32(v) for v in rgb)
    m = max(r, g, b)
    if m <= 0:
        return (f32(0), f32(0), f32(-1e10)), f32(0)
    scale = f32(f32(2) * m)
    return rgb_to_spectrum(f32(r / scale), f32(g / scale), f32(b / scale)), scale
