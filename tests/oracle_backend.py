"""Loader for the CPU oracle (oracle/libhikari_oracle.so) — TEST INFRASTRUCTURE.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os

import numpy as np

from hikari_jl_b200 import _abi as A
from hikari_jl_b200.host import Backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "libhikari_oracle.so")
_lib = None
VP = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        A.bind_common(L, "ok_")
        def f(name, args, res=C.c_int32):
            fn = getattr(L, name); fn.argtypes = args; fn.restype = res
        f("ok_trace_closest", [VP, A.c_fp, C.c_uint64, A.c_fp, C.c_int32])
        f("ok_set_brute_force", [VP, C.c_int32])
        f("ok_set_row_subset", [VP, C.c_int32, C.c_int32])
        f("ok_test_detmath", [C.c_int32, A.c_fp, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_num_threads", [])
        f("ok_set_num_threads", [C.c_int32])
        f("ok_rays_traced", [VP], C.c_uint64)
        f("ok_read_pixel_L", [VP, A.c_fp, A.c_fp, A.c_fp, A.c_fp])
        f("ok_test_sobol", [VP, A.c_i32p, C.c_uint64, C.c_int32, C.c_int32, C.c_uint32, A.c_fp, A.c_fp])
        f("ok_test_hashes", [A.c_fp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), A.c_fp])
        f("ok_murmur64a", [A.c_u8p, C.c_uint64, C.c_uint64], C.c_uint64)
        f("ok_pcg32_stream", [C.c_uint64, C.c_uint64, A.c_u32p, C.c_int32])
        f("ok_sobol_raw", [VP, C.c_int64, C.c_int32, A.c_u32p])
        f("ok_test_wavelengths", [A.c_fp, C.c_uint64, A.c_fp, A.c_fp])
        f("ok_test_uplift", [VP, C.c_int32, A.c_fp, A.c_fp, C.c_uint64, A.c_fp, A.c_fp])
        f("ok_test_spectral_to_rgb", [VP, A.c_fp, A.c_fp, A.c_fp, C.c_uint64, A.c_fp, A.c_fp])
        f("ok_test_filter", [VP, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_camera_rays", [VP, C.c_int32, A.c_fp])
        f("ok_test_bsdf", [VP, C.c_uint32, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_lights", [VP, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_escaped", [VP, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_delta_tracking", [VP, C.c_uint32, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_density", [VP, C.c_uint32, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_test_ratio_tracking", [VP, C.c_uint32, A.c_fp, C.c_uint64, A.c_fp])
        f("ok_fresnel_dielectric", [C.c_float, C.c_float], C.c_float)
        f("ok_fr_complex", [C.c_float, C.c_float, C.c_float], C.c_float)
        f("ok_write_accum", [VP, A.c_fp, A.c_fp])
        f("ok_test_set_aux_depth", [VP, A.c_fp])
        f("ok_mix_hash_float", [A.c_fp, A.c_fp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32], C.c_float)
        _lib = L
    return _lib


class OracleBackend(Backend):
    """The checker behind the product's host-side flattening (Backend.upload_scene / set_params / read_film): same structs, the
    ok_* entry points of oracle/libhikari_oracle.so instead of hk_*."""
    prefix = "ok_"
    device_majorant = False       # the checker is handed the host-built (numpy) majorant grids

    def __init__(self):
        self.lib = lib()
        self.ctx = C.c_void_p()
        rc = self.lib.ok_create(C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError(f"ok_create failed with status {rc}")
        self._keep = []

    def _last_error(self):
        return ""


def make_backend():
    return OracleBackend()


def fp(a):
    return a.ctypes.data_as(A.c_fp)
