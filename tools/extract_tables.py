#!/usr/bin/env python3
"""Extract the constant numeric DATA tables the VolPath hot path needs from the
reference checkout into data/hikari_tables.npz.

Only numbers are extracted (published pbrt-v4 / CIE data); no reference code is copied.
Run in the build container (needs /root/reference):  python tools/extract_tables.py

Tables and where they come from (reference file:line):
  sobol_matrices  u32[1024*52]  src/sampler/sobol_matrices.jl:18-6675  (Joe-Kuo / pbrt-v4 SobolMatrices32)
  cie_x/y/z       f32[471]      src/spectral/color.jl:53,151,249       (CIE 1931, 1 nm, 360-830 nm)
  d65_values      f32[107]      src/spectral/uplift.jl:412-429         (CIE D65, 5 nm, 300-830 nm)
  gen_cie_x/y/z   f64[95]       src/spectral/rgb2spec_gen.jl:20,42,64  (CIE 1931, 5 nm, for the table generator)
  gen_d65         f64[95]       src/spectral/rgb2spec_gen.jl:88        (D65 5 nm normalised by CIE_D65_NORM)
  <metal>_eta/_k  f32[2,N]      src/spectral/metal-spectra.jl          (lambda row, value row)
"""
import re, sys, os
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "hikari_tables.npz")

def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()

def block_after(src, start_pat, open_ch, close_ch):
    m = re.search(start_pat, src)
    assert m, start_pat
    i = src.index(open_ch, m.end() - 1)
    depth = 0
    for j in range(i, len(src)):
        if src[j] == open_ch: depth += 1
        elif src[j] == close_ch:
            depth -= 1
            if depth == 0:
                return src[i + 1:j]
    raise ValueError(start_pat)

def strip_comments(s):
    return re.sub(r"#.*", "", s)

def floats(s):
    s = strip_comments(s)
    s = re.sub(r"f0\b", "", s)
    s = re.sub(r"(\d)f(-?\d+)", r"\1e\2", s)
    return [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", s)]

out = {}
# Sobol
src = read("src/sampler/sobol_matrices.jl")
blk = block_after(src, r"const SobolMatrices32 = UInt32\[", "[", "]")
vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", blk)]
assert len(vals) == 1024 * 52, len(vals)
out["sobol_matrices"] = np.array(vals, dtype=np.uint32)

# CIE 1 nm
src = read("src/spectral/color.jl")
for nm in ("CIE_X", "CIE_Y", "CIE_Z"):
    v = floats(block_after(src, rf"const {nm} = Float32\[", "[", "]"))
    assert len(v) == 471, (nm, len(v))
    out[nm.lower()] = np.array(v, dtype=np.float32)

# D65 5 nm
src = read("src/spectral/uplift.jl")
v = floats(block_after(src, r"const D65_ILLUMINANT_VALUES = \(", "(", ")"))
assert len(v) == 107, len(v)
out["d65_values"] = np.array(v, dtype=np.float32)

# generator tables (Float64)
src = read("src/spectral/rgb2spec_gen.jl")
for nm in ("CIE_X", "CIE_Y", "CIE_Z"):
    v = floats(block_after(src, rf"const {nm} = Float64\[", "[", "]"))
    assert len(v) == 95, (nm, len(v))
    out["gen_" + nm.lower()] = np.array(v, dtype=np.float64)
m = re.search(r"const CIE_D65_NORM = ([\d.]+)", src)
norm = float(m.group(1))
blk = block_after(src, r"const CIE_D65 = Float64\[", "[", "]")
blk = strip_comments(blk)
# entries may be written as "x / CIE_D65_NORM" or as a broadcast after the literal
raw = floats(blk.replace("CIE_D65_NORM", ""))
print("gen_d65 raw count", len(raw), "norm", norm)
tail = src[src.index("const CIE_D65 = Float64["):][:6000]
out["gen_d65_raw"] = np.array(raw, dtype=np.float64)
out["gen_d65_norm"] = np.array([norm], dtype=np.float64)

# metals
src = read("src/spectral/metal-spectra.jl")
for m in re.finditer(r"const (\w+)_(ETA|K)_SPECTRUM = from_interleaved\(PiecewiseLinearSpectrum\{(\d+)\}, \(", src):
    name, kind, n = m.group(1).lower(), m.group(2).lower(), int(m.group(3))
    blk = block_after(src[m.start():], r"from_interleaved\(PiecewiseLinearSpectrum\{\d+\}, \(", "(", ")")
    v = floats(blk)
    # first paren group captured is the outer call; take the inner tuple numbers only
    v = [x for x in v]
    # drop the leading N of PiecewiseLinearSpectrum{N} if it slipped in
    if len(v) == 2 * n + 1: v = v[1:]
    assert len(v) == 2 * n, (name, kind, len(v), n)
    a = np.array(v, dtype=np.float32).reshape(n, 2).T.copy()
    out[f"{name}_{kind}"] = a

np.savez_compressed(OUT, **out)
print("wrote", OUT, {k: v.shape for k, v in out.items()})
