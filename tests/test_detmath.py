"""csrc/hk_detmath.h — the ONE f32 libm compiled into both the CUDA library and the oracle.
CPU: accuracy of the host evaluation against double-precision libm.  GPU: the device evaluation of the same header
equals the host evaluation bit for bit (which is what makes the hashed-RNG stages of the path reproducible)."""
import ctypes as C

import numpy as np
import pytest

import oracle_backend
from hikari_jl_b200 import _abi as A
from util import gpu_available

f32 = np.float32
FN = {"expf": 0, "logf": 1, "sinf": 2, "cosf": 3, "coshf": 4, "atanhf": 5, "powf": 6, "log1pf": 7}


def fp(a):
    return a.ctypes.data_as(A.c_fp)


def ok_eval(fn, x, y=None):
    L = oracle_backend.lib()
    L.ok_test_detmath.argtypes = [C.c_int32, A.c_fp, A.c_fp, C.c_uint64, A.c_fp]
    out = np.zeros(len(x), dtype=f32)
    L.ok_test_detmath(FN[fn], fp(x), fp(y) if y is not None else None, len(x), fp(out))
    return out


def ulp_err(got, want64):
    want32 = want64.astype(f32)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    ulp = np.maximum(ulp, np.float64(np.finfo(f32).smallest_subnormal))
    return np.abs(got.astype(np.float64) - want64) / ulp


def samples(lo, hi, n, seed):
    rng = np.random.RandomState(seed)
    return np.concatenate([rng.uniform(lo, hi, n), np.linspace(lo, hi, 4097)]).astype(f32)


CASES = [
    # fn, inputs, double reference, max ulp
    ("expf", lambda: samples(-87.0, 88.0, 400000, 1), np.exp, 1.5),
    ("expf", lambda: samples(-1.0, 1.0, 200000, 2), np.exp, 1.5),
    ("logf", lambda: np.exp(samples(-87.0, 88.0, 400000, 3).astype(np.float64)).astype(f32), np.log, 1.5),
    ("logf", lambda: samples(0.5, 2.0, 400000, 4), np.log, 1.5),
    ("sinf", lambda: samples(-7.0, 7.0, 400000, 5), np.sin, 2.0),
    ("cosf", lambda: samples(-7.0, 7.0, 400000, 6), np.cos, 2.0),
    ("sinf", lambda: samples(-30000.0, 30000.0, 400000, 7), np.sin, 2.5),
    ("cosf", lambda: samples(-30000.0, 30000.0, 400000, 8), np.cos, 2.5),
    ("sinf", lambda: samples(-1.0e7, 1.0e7, 200000, 9), np.sin, 2.0),
    ("coshf", lambda: samples(-10.0, 10.0, 200000, 10), np.cosh, 3.0),
    ("atanhf", lambda: samples(-0.98, 0.98, 400000, 11), np.arctanh, 3.0),
    ("log1pf", lambda: samples(-0.9, 100.0, 200000, 12), np.log1p, 3.0),
]


@pytest.mark.parametrize("fn,gen,ref,tol", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_accuracy_vs_double_libm(fn, gen, ref, tol):
    x = gen()
    got = ok_eval(fn, x)
    err = ulp_err(got, ref(x.astype(np.float64)))
    assert err.max() <= tol, f"{fn}: max error {err.max():.3f} ulp at x = {x[err.argmax()]!r}"


def test_powf_accuracy_and_special_cases():
    rng = np.random.RandomState(20)
    x = rng.uniform(0.0, 4.0, 300000).astype(f32); y = rng.uniform(-8.0, 130.0, 300000).astype(f32)
    got = ok_eval("powf", x, y)
    with np.errstate(all="ignore"):
        want = np.power(x.astype(np.float64), y.astype(np.float64))
    fin = np.isfinite(want) & (want < 3.0e38) & (want > 1.0e-37)
    assert ulp_err(got[fin], want[fin]).max() <= 1.0
    sx = np.array([0.0, 0.0, -0.0, 2.0, -2.0, -2.0, -2.0, 1.0, np.nan, 5.0, np.inf, 0.5, 0.5, -8.0], dtype=f32)
    sy = np.array([0.0, 2.0, 3.0, np.inf, 2.0, 3.0, 0.5, np.nan, 0.0, -np.inf, -1.0, np.inf, -np.inf, 1.0 / 3.0], dtype=f32)
    g = ok_eval("powf", sx, sy)
    with np.errstate(all="ignore"):
        w = np.power(sx.astype(np.float64), sy.astype(np.float64)).astype(f32)
    assert np.array_equal(np.isnan(g), np.isnan(w))
    assert np.array_equal(g[~np.isnan(g)], w[~np.isnan(w)])
    assert np.array_equal(np.signbit(g[~np.isnan(g)]), np.signbit(w[~np.isnan(w)]))


def test_special_values():
    inf, nan = f32(np.inf), f32(np.nan)
    e = ok_eval("expf", np.array([0.0, -inf, inf, nan, 89.0, -104.5, -90.0], dtype=f32))
    assert e[0] == 1.0 and e[1] == 0.0 and np.isinf(e[2]) and np.isnan(e[3]) and np.isinf(e[4]) and e[5] == 0.0
    assert 0.0 < e[6] < 1.2e-38 and abs(float(e[6]) / np.exp(-90.0) - 1.0) < 1e-4          # denormal result
    l = ok_eval("logf", np.array([1.0, 0.0, -0.0, -1.0, inf, nan, 1.0e-40], dtype=f32))
    assert l[0] == 0.0 and l[1] == -inf and l[2] == -inf and np.isnan(l[3]) and l[4] == inf and np.isnan(l[5])
    assert abs(float(l[6]) - np.log(float(f32(1.0e-40)))) < 1e-5
    s = ok_eval("sinf", np.array([0.0, inf, nan, 1.0e30], dtype=f32))
    assert s[0] == 0.0 and np.isnan(s[1]) and np.isnan(s[2]) and abs(s[3]) <= 1.0
    a = ok_eval("atanhf", np.array([0.0, 1.0, -1.0, 1.5], dtype=f32))
    assert a[0] == 0.0 and a[1] == inf and a[2] == -inf and np.isnan(a[3])


@pytest.mark.gpu
def test_device_equals_host_bit_for_bit():
    if not gpu_available():
        pytest.skip("needs a CUDA device")
    from hikari_jl_b200.host import Backend
    be = Backend()
    be.lib.hk_test_detmath.argtypes = [C.c_void_p, C.c_int32, A.c_fp, A.c_fp, C.c_uint64, A.c_fp]
    rng = np.random.RandomState(99)
    n = 2_000_000
    # every exponent / sign / mantissa pattern, plus the dense ranges the path uses
    bits = rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(f32)
    dense = {
        "expf": rng.uniform(-110, 95, n), "logf": np.abs(bits), "sinf": rng.uniform(-40000, 40000, n), "cosf": rng.uniform(-40000, 40000, n),
        "coshf": rng.uniform(-95, 95, n), "atanhf": rng.uniform(-1.01, 1.01, n), "log1pf": rng.uniform(-1.1, 50, n),
    }
    for fn, k in FN.items():
        if fn == "powf":
            continue
        for x in (bits, dense[fn].astype(f32)):
            host = ok_eval(fn, x)
            dev = np.zeros(n, dtype=f32)
            rc = be.lib.hk_test_detmath(be.ctx, k, fp(x), None, n, fp(dev))
            assert rc == 0, be.last_error()
            same = host.view(np.uint32) == dev.view(np.uint32)
            both_nan = np.isnan(host) & np.isnan(dev)
            bad = ~(same | both_nan)
            assert not bad.any(), f"{fn}: {bad.sum()} of {n} differ, first x = {x[bad][0]!r}: host {host[bad][0]!r} device {dev[bad][0]!r}"
    for x, y in ((rng.uniform(0, 8, n).astype(f32), rng.uniform(-40, 140, n).astype(f32)), (bits, np.roll(bits, 1))):
        host = ok_eval("powf", x, y)
        dev = np.zeros(n, dtype=f32)
        assert be.lib.hk_test_detmath(be.ctx, 6, fp(x), fp(y), n, fp(dev)) == 0
        bad = ~((host.view(np.uint32) == dev.view(np.uint32)) | (np.isnan(host) & np.isnan(dev)))
        assert not bad.any(), f"powf: {bad.sum()} differ, first (x, y) = ({x[bad][0]!r}, {y[bad][0]!r})"
    be.close()
