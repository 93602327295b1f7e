"""CPU tests: the C-ABI library loads without a GPU and exports every symbol include/*.h declares; compute
entry points fail loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

from hikari_jl_b200 import _abi as A
from util import gpu_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hk_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_loads_and_exports_every_declared_symbol():
    lib = A.load_library()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    for s in A.HK_SYMBOLS:
        assert s in syms, f"{s} listed in _abi.HK_SYMBOLS but not declared in the header"
    assert lib.hk_abi_version() == 4


def test_struct_sizes_match_header_layout():
    # sizes computed from the header by hand; a drift between ctypes and C would corrupt every upload
    assert C.sizeof(A.HkMaterial) == 4 + 4 + 12 + 12 + 16 + 32 + 8 + 8 + 16 + 32
    assert C.sizeof(A.HkLightBVHNode) == 64
    assert C.sizeof(A.HkMediumInterface) == 12
    assert C.sizeof(A.HkMedium) == 416      # ABI v4: + nanovdb_index_min / _max
    assert C.sizeof(A.HkRenderParams) == 44
    assert C.sizeof(A.HkLight) == 4 * (2 + 1 + 3 + 3 + 1 + 3 + 3 + 2 + 16 + 9 + 3 + 1 + 6 + 2)


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    lib = A.load_library()
    ctx = C.c_void_p()
    rc = lib.hk_create(0, C.byref(ctx))
    assert rc == -3, "hk_create must fail with HK_ERR_NO_DEVICE when no CUDA device is present"
    from hikari_jl_b200.host import Backend
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Backend()


def _build_c_client(tmp_path):
    """gcc (C, not C++) compiles tests/c/abi_client.c against include/hikari_cuda.h and links libhikari_cuda.so"""
    import subprocess
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(A.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-include", os.path.join(ROOT, "include", "hikari_cuda_testing.h"),
           os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", exe, "-L", libdir, "-lhikari_cuda", "-lm", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _write_tables(path):
    import numpy as np
    from hikari_jl_b200 import tables as T
    t = T.load_tables()
    scale, coeffs = T.get_srgb_table()
    with open(path, "wb") as f:
        f.write(np.int32(len(scale)).tobytes())
        for k in ("sobol_matrices", "cie_x", "cie_y", "cie_z", "d65_values"):
            f.write(np.ascontiguousarray(t[k]).tobytes())
        f.write(np.ascontiguousarray(scale, dtype=np.float32).tobytes()); f.write(np.ascontiguousarray(coeffs, dtype=np.float32).tobytes())


@pytest.mark.skipif(gpu_available(), reason="the no-GPU half of the C client")
def test_c_client_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    import subprocess
    A.load_library()
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("no-device abi=4"), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_client_renders_through_the_c_abi(tmp_path):
    """A C program (no Python, no ctypes) drives create -> uploads -> render -> host and device read-out -> destroy."""
    import subprocess
    A.load_library()
    exe = _build_c_client(tmp_path)
    tab = str(tmp_path / "tables.bin")
    _write_tables(tab)
    r = subprocess.run([exe, tab], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "dev_same=1" in r.stdout, (r.returncode, r.stdout, r.stderr)


def _julia_structs():
    """struct mirrors of julia/HikariCUDA.jl: {name: [(field, julia type), ...]} in declaration order"""
    lines = open(os.path.join(ROOT, "julia", "HikariCUDA.jl")).read().splitlines()
    out, k = {}, 0
    while k < len(lines):
        one = re.match(r"^struct (Hk\w+);(.*); end\s*(#.*)?$", lines[k])
        multi = re.match(r"^(?:Base\.@kwdef )?struct (Hk\w+)\s*(#.*)?$", lines[k])
        if one:
            name, body = one.group(1), one.group(2)
        elif multi:
            name, body = multi.group(1), ""
            k += 1
            while lines[k] != "end":
                body += re.sub(r"#.*", "", lines[k]) + "\n"
                k += 1
        else:
            k += 1
            continue
        fields = []
        for part in re.split(r"[;\n]", body):
            part = part.split(" = ")[0].strip()
            if "::" in part:
                f, t = part.split("::")
                fields.append((f.strip(), t.strip()))
        out[name] = fields
        k += 1
    return out


def _julia_type_of(ct):
    """the Julia spelling of a ctypes field type"""
    scalars = {C.c_int32: "Int32", C.c_uint32: "UInt32", C.c_float: "Float32", C.c_uint64: "UInt64", C.c_int64: "Int64", C.c_uint8: "UInt8"}
    if ct in scalars:
        return scalars[ct]
    if ct is C.c_void_p:
        return "Ptr{Cvoid}"
    if hasattr(ct, "_length_"):
        return "NTuple{%d,%s}" % (ct._length_, _julia_type_of(ct._type_))
    if hasattr(ct, "contents") or hasattr(ct, "_type_"):
        inner = ct._type_
        return "Ptr{%s}" % (inner.__name__ if issubclass(inner, C.Structure) else _julia_type_of(inner))
    raise AssertionError(ct)


def test_julia_struct_mirrors_match_the_ctypes_mirrors():
    """julia/HikariCUDA.jl cannot be executed here, so its struct mirrors are checked textually: every Hk* struct it declares
    has the fields of the ctypes mirror (which the C client test and every GPU test exercise), same names, order and types."""
    js = _julia_structs()
    assert len(js) >= 18, sorted(js)
    for name, fields in js.items():
        ct = getattr(A, name)
        want = [(f, _julia_type_of(t)) for f, t in ct._fields_]
        assert fields == want, f"{name}: julia {fields} != ctypes {want}"


def test_ctypes_mirrors_match_the_header_field_by_field(tmp_path):
    """Every struct of include/hikari_cuda.h against its ctypes mirror: a generated C program prints sizeof and offsetof of every field
    the mirror names (gcc against the real header), and the numbers must equal ctypes' -- names, order, padding and sizes."""
    import subprocess
    structs = [n for n in dir(A) if n.startswith("Hk") and isinstance(getattr(A, n), type) and issubclass(getattr(A, n), C.Structure)]
    assert len(structs) >= 18
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "hikari_cuda.h"', 'int main(void) {']
    want = []
    for n in sorted(structs):
        ct = getattr(A, n)
        src.append(f'  printf("%zu\\n", sizeof({n}));')
        want.append((n, "sizeof", C.sizeof(ct)))
        for f, _ in ct._fields_:
            src.append(f'  printf("%zu\\n", offsetof({n}, {f}));')
            want.append((n, f, getattr(ct, f).offset))
    src += ['  return 0;', '}']
    cfile, exe = tmp_path / "layout.c", tmp_path / "layout"
    cfile.write_text("\n".join(src))
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(cfile), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert len(got) == len(want)
    bad = [(n, f, w, g) for (n, f, w), g in zip(want, got) if w != g]
    assert not bad, f"ctypes vs C layout (struct, field, ctypes, C): {bad[:8]}"


def _split_top_level(s):
    """split on commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_ccalls_match_the_header_prototypes():
    """Every ccall of julia/HikariCUDA.jl names an entry point the header declares, with as many argument types as the C prototype
    has parameters, Int32 as the return type of the status-returning calls, and pointer / scalar kinds that agree position by
    position (Ptr / Ref / Cstring for C pointers; Int32 / UInt32 / UInt64 / Float32 for the scalars)."""
    hdr = open(os.path.join(ROOT, "include", "hikari_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*([A-Za-z_][\w \*]*?)\b(hk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else _split_top_level(params)
        protos[name] = (ret, ["*" in p for p in plist])
    jl = open(os.path.join(ROOT, "julia", "HikariCUDA.jl")).read()
    calls = re.findall(r"ccall\(\(:(hk_[a-z0-9_]+), lib\), (\w+), \(((?:[^()]|\([^()]*\))*?)\)", jl)
    assert len(calls) >= 25, len(calls)
    for name, ret, types in calls:
        assert name in protos, f"{name}: not declared in include/hikari_cuda.h"
        cret, cptr = protos[name]
        tl = [t for t in _split_top_level(types) if t]
        assert len(tl) == len(cptr), f"{name}: ccall passes {len(tl)} arguments {tl}, the header declares {len(cptr)}"
        assert (ret == "Int32") == (cret == "int32_t"), f"{name}: return type {ret} vs {cret}"
        for k, (t, is_ptr) in enumerate(zip(tl, cptr)):
            j_ptr = t.startswith(("Ptr{", "Ref{")) or t == "Cstring"
            assert j_ptr == is_ptr, f"{name}: argument {k + 1} is {t} in the ccall, {'a pointer' if is_ptr else 'a scalar'} in the header"
