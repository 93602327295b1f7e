"""Shared helpers for the parity tests."""
import ctypes as C
import numpy as np

from hikari_jl_b200 import _abi as A
from hikari_jl_b200.host import Backend, Film, VolPath, FilterSamplerData, GaussianFilter
import oracle_backend

f32 = np.float32


def fp(a):
    return a.ctypes.data_as(A.c_fp)


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class Pair:
    """The SAME flattened scene uploaded to the CUDA library (hk_) and to the CPU oracle (ok_)."""

    def __init__(self, scene=None, film=None, camera=None, vp=None, need_gpu=True, sample_batch=1):
        self.cu = Backend() if need_gpu else None
        self.ok = oracle_backend.make_backend()
        self.lib = self.cu.lib if self.cu else None
        self.olib = oracle_backend.lib()
        for b in (self.cu, self.ok):
            if b is None:
                continue
            b.upload_tables()
            if scene is not None:
                b.upload_scene(scene)
            fsd = vp.filter_sampler_data if vp is not None else FilterSamplerData(GaussianFilter())
            b.set_filter(fsd)
            if film is not None:
                v = vp if vp is not None else VolPath(samples=1, max_depth=5)
                b.set_params(v, film.resolution[0], film.resolution[1], sample_batch if b is self.cu else 1)
            if camera is not None:
                b.set_camera(camera)

    def close(self):
        if self.cu:
            self.cu.close()
        self.ok.close()


def image_close(a, b, atol=1e-3, rtol=2e-2):
    """SURVEY 8c image tolerance: |a-b| <= 1e-3 + 2e-2*max(a,b); returns (fraction within, relative RMSE)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))
    rrmse = np.sqrt(np.mean((a - b) ** 2)) / max(1e-12, np.sqrt(np.mean(b ** 2)))
    return float(ok.mean()), float(rrmse)


def random_rays(n, seed, lo=-3.0, hi=3.0, tmax=np.inf):
    rng = np.random.RandomState(seed)
    o = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8), dtype=f32)
    r[:, 0:3] = o; r[:, 3:6] = d; r[:, 6] = tmax; r[:, 7] = 0
    return r


def trace_both(pair, rays, brute=False):
    n = len(rays)
    h_cu = np.zeros((n, 4), dtype=f32); h_ok = np.zeros((n, 4), dtype=f32)
    rc = pair.lib.hk_trace_closest(pair.cu.ctx, fp(rays), n, fp(h_cu))
    assert rc == 0, pair.lib.hk_last_error(pair.cu.ctx)
    pair.olib.ok_trace_closest(pair.ok.ctx, fp(rays), n, fp(h_ok), 1 if brute else 0)
    return h_cu, h_ok


def assert_bits_equal(a, b, what, allow=0):
    """Bit-for-bit equality of two f32 arrays (NaN == NaN whatever the payload); `allow` = rows that may differ."""
    a = np.ascontiguousarray(a, dtype=f32); b = np.ascontiguousarray(b, dtype=f32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    rows = same.reshape(len(a), -1).all(axis=1)
    bad = int((~rows).sum())
    if bad > allow:
        i = int(np.argmin(rows))
        raise AssertionError(f"{what}: {bad} of {len(a)} records differ bitwise (allowed {allow}); first at {i}: {a.reshape(len(a), -1)[i]} vs {b.reshape(len(b), -1)[i]}")
