"""ctypes mirror of include/hikari_cuda.h (struct layouts + prototypes) and the loader of
libhikari_cuda.so.  There is NO CPU fallback: if the CUDA library cannot be loaded, or no CUDA device
is usable, every compute entry point raises.
"""
import ctypes as C
import os

c_f = C.c_float
c_fp = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)

HK_MAT_MATTE, HK_MAT_MIRROR, HK_MAT_GLASS, HK_MAT_CONDUCTOR = 1, 2, 3, 4
HK_MAT_COATED_DIFFUSE, HK_MAT_THIN_DIELECTRIC, HK_MAT_DIFFUSE_TRANSMISSION = 5, 6, 7
HK_MAT_MIX = 8
HK_MAT_COATED_CONDUCTOR = 9
HK_MAT_COATED_DIFFUSE_TRANSMISSION = 10
HK_MATFLAG_REMAP_ROUGHNESS, HK_MATFLAG_SPECTRAL_ETA_K, HK_MATFLAG_USE_ETA_K, HK_MATFLAG_VERTEX_COLORS = 1, 2, 4, 8
HK_LIGHT_POINT, HK_LIGHT_SPOT, HK_LIGHT_DIRECTIONAL, HK_LIGHT_SUN = 1, 2, 3, 4
HK_LIGHT_ENVIRONMENT, HK_LIGHT_AMBIENT, HK_LIGHT_DIFFUSE_AREA = 5, 6, 7
HK_SPECTRUM_RGB, HK_SPECTRUM_ILLUMINANT = 0, 1
HK_MEDIUM_HOMOGENEOUS, HK_MEDIUM_GRID, HK_MEDIUM_NANOVDB, HK_MEDIUM_RGBGRID = 1, 2, 3, 4


class HkTables(C.Structure):
    _fields_ = [("sobol_matrices", c_u32p), ("cie_x", c_fp), ("cie_y", c_fp), ("cie_z", c_fp), ("d65", c_fp),
                ("rgb2spec_res", C.c_int32), ("rgb2spec_scale", c_fp), ("rgb2spec_coeffs", c_fp)]


class HkPostprocess(C.Structure):
    _fields_ = [("exposure", c_f), ("tonemap_mode", C.c_int32), ("inv_gamma", c_f), ("apply_gamma", C.c_int32),
                ("white_point", c_f), ("imaging_ratio", c_f), ("apply_wb", C.c_int32), ("wb", c_f * 9),
                ("mask_escaped", C.c_int32), ("background", c_f * 3)]


class HkDenoiseConfig(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("sigma_color", c_f), ("sigma_normal", c_f), ("sigma_depth", c_f),
                ("use_variance", C.c_int32)]


class HkMesh(C.Structure):
    _fields_ = [("first_tri", C.c_uint32), ("n_tris", C.c_uint32)]


class HkInstance(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("medium_interface_idx", C.c_uint32), ("object_to_world", c_f * 12), ("world_to_object", c_f * 12)]


class HkGeometry(C.Structure):
    _fields_ = [("positions", c_fp), ("normals", c_fp), ("tangents", c_fp), ("uvs", c_fp), ("indices", c_u32p),
                ("tri_meta", c_u32p), ("n_verts", C.c_uint32), ("n_tris", C.c_uint32),
                ("meshes", C.POINTER(HkMesh)), ("n_meshes", C.c_uint32), ("instances", C.POINTER(HkInstance)), ("n_instances", C.c_uint32)]


class HkMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("flags", C.c_uint32), ("rgb0", c_f * 3), ("rgb1", c_f * 3), ("rgb2", c_f * 4), ("f", c_f * 8),
                ("spec", C.c_int32 * 2), ("ival", C.c_int32 * 2), ("tex", C.c_int32 * 4), ("ftex", C.c_int32 * 8)]


class HkTexture(C.Structure):
    _fields_ = [("rgb", c_fp), ("h", C.c_int32), ("w", C.c_int32), ("alpha", c_fp)]


class HkMediumInterface(C.Structure):
    _fields_ = [("material", C.c_uint32), ("inside", C.c_uint32), ("outside", C.c_uint32)]


class HkSpectra(C.Structure):
    _fields_ = [("lambdas", c_fp), ("values", c_fp), ("offsets", c_u32p), ("n_spectra", C.c_uint32)]


class HkLight(C.Structure):
    _fields_ = [("type", C.c_int32), ("spectrum_kind", C.c_int32), ("scale", c_f), ("rgb", c_f * 3), ("poly", c_f * 3),
                ("illum_scale", c_f), ("position", c_f * 3), ("direction", c_f * 3), ("cos_total_width", c_f),
                ("cos_falloff_start", c_f), ("world_to_light", c_f * 16), ("v", c_f * 9), ("normal", c_f * 3),
                ("area", c_f), ("uv", c_f * 6), ("two_sided", C.c_int32), ("env_map", C.c_int32)]


class HkEnvMap(C.Structure):
    _fields_ = [("rgb", c_fp), ("w", C.c_int32), ("h", C.c_int32), ("rotation", c_f * 9), ("scale_rgb", c_f * 3),
                ("conditional_func", c_fp), ("conditional_cdf", c_fp), ("conditional_func_int", c_fp),
                ("marginal_func", c_fp), ("marginal_cdf", c_fp), ("marginal_func_int", c_f),
                ("nu", C.c_int32), ("nv", C.c_int32)]


class HkLightBVHNode(C.Structure):
    _fields_ = [("bounds_min", c_f * 3), ("bounds_max", c_f * 3), ("w", c_f * 3), ("phi", c_f), ("cos_theta_o", c_f),
                ("cos_theta_e", c_f), ("two_sided", C.c_uint32), ("child1_or_light_idx", C.c_uint32),
                ("is_leaf", C.c_uint32), ("_pad", C.c_uint32)]


class HkLightSampler(C.Structure):
    _fields_ = [("nodes", C.POINTER(HkLightBVHNode)), ("n_nodes", C.c_uint32), ("light_to_bit_trail", c_u32p),
                ("infinite_light_indices", c_i32p), ("n_infinite", C.c_uint32), ("n_bvh_lights", C.c_uint32)]


class HkMedium(C.Structure):
    _fields_ = [("type", C.c_int32), ("sigma_a_rgb", c_f * 3), ("sigma_s_rgb", c_f * 3), ("Le_rgb", c_f * 3),
                ("scale", c_f), ("g", c_f), ("bounds_min", c_f * 3), ("bounds_max", c_f * 3),
                ("render_from_medium", c_f * 16), ("medium_from_render", c_f * 16), ("density_res", C.c_int32 * 3),
                ("density", c_fp), ("majorant_res", C.c_int32 * 3), ("majorant", c_fp), ("nanovdb_buf", c_u8p),
                ("nanovdb_bytes", C.c_uint64), ("nanovdb_inv_mat", c_f * 9), ("nanovdb_vec", c_f * 3),
                ("nanovdb_root_offset", C.c_uint64), ("nanovdb_upper_offset", C.c_uint64),
                ("nanovdb_lower_offset", C.c_uint64), ("nanovdb_leaf_offset", C.c_uint64),
                ("nanovdb_root_tiles", C.c_int32), ("nanovdb_upper_count", C.c_int32),
                ("nanovdb_lower_count", C.c_int32), ("nanovdb_leaf_count", C.c_int32),
                ("rgb_sigma_a", c_fp), ("rgb_sigma_s", c_fp), ("rgb_Le", c_fp), ("Le_scale", c_f),
                ("nanovdb_index_min", C.c_int32 * 3), ("nanovdb_index_max", C.c_int32 * 3)]


class HkCamera(C.Structure):
    _fields_ = [("raster_to_camera", c_f * 16), ("camera_to_world", c_f * 16), ("lens_radius", c_f),
                ("focal_distance", c_f), ("shutter_open", c_f), ("shutter_close", c_f), ("dx_camera", c_f * 3),
                ("dy_camera", c_f * 3)]


class HkFilter(C.Structure):
    _fields_ = [("type", C.c_int32), ("radius", c_f * 2), ("nx", C.c_int32), ("ny", C.c_int32), ("func", c_fp),
                ("marginal_cdf", c_fp), ("marginal_func", c_fp), ("conditional_cdf", c_fp), ("domain_min", c_f * 2),
                ("domain_max", c_f * 2), ("func_integral", c_f)]


class HkRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32),
                ("samples_per_pixel", C.c_int32), ("regularize", C.c_int32), ("max_component_value", c_f),
                ("sampler_seed", C.c_uint32), ("sobol_log2_spp", C.c_int32), ("sobol_n_base4_digits", C.c_int32),
                ("material_coherence", C.c_int32), ("sample_batch", C.c_int32)]


class HkStats(C.Structure):
    _fields_ = [("rays_traced", C.c_uint64), ("samples_rendered", C.c_uint64), ("queue_overflows", C.c_uint64),
                ("bvh_nodes", C.c_uint64), ("bvh_bytes", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("last_render_ms", c_f), ("last_trace_ms", c_f), ("path_vertices", C.c_uint64)]


# every symbol include/hikari_cuda.h declares (checked by tests/test_abi.py)
HK_SYMBOLS = [
    "hk_abi_version", "hk_create", "hk_destroy", "hk_last_error", "hk_upload_tables", "hk_upload_geometry",
    "hk_upload_spectra", "hk_upload_textures", "hk_upload_materials", "hk_update_material", "hk_bounce_profile", "hk_upload_envmaps", "hk_upload_lights", "hk_upload_media", "hk_update_medium", "hk_read_majorant", "hk_read_nanovdb",
    "hk_set_camera", "hk_set_filter", "hk_set_params", "hk_clear", "hk_render_samples", "hk_render_samples_strided",
    "hk_read_film", "hk_read_film_async", "hk_read_film_wait", "hk_read_film_dev", "hk_set_stream", "hk_postprocess", "hk_postprocess_dev", "hk_fill_aux_buffers", "hk_read_aux_buffers", "hk_denoise", "hk_film_accum_dev", "hk_read_accum", "hk_write_accum", "hk_trace_closest",
    "hk_trace_closest_dev", "hk_trace_any", "hk_stats", "hk_synchronize", "hk_dev_alloc", "hk_dev_free",
    "hk_dev_upload", "hk_dev_download", "hk_set_profiling", "hk_stage_times", "hk_pinned_alloc", "hk_pinned_free",
]

_VP = C.c_void_p


def bind_common(lib, p):
    """Set prototypes of the scene/render entry points shared by hk_* (CUDA) and ok_* (oracle)."""
    def f(name, args, res=C.c_int32):
        fn = getattr(lib, p + name)
        fn.argtypes = args
        fn.restype = res
        return fn
    f("create", [C.c_int32, C.POINTER(_VP)] if p == "hk_" else [C.POINTER(_VP)])
    f("destroy", [_VP])
    f("upload_tables", [_VP, C.POINTER(HkTables)])
    f("upload_geometry", [_VP, C.POINTER(HkGeometry)])
    f("upload_spectra", [_VP, C.POINTER(HkSpectra)])
    f("upload_materials", [_VP, C.POINTER(HkMaterial), C.c_uint32, C.POINTER(HkMediumInterface), C.c_uint32])
    f("update_material", [_VP, C.c_uint32, C.POINTER(HkMaterial)])
    f("postprocess", [_VP, C.POINTER(HkPostprocess), c_fp])
    f("fill_aux_buffers", [_VP, C.c_int32])
    f("read_aux_buffers", [_VP, c_fp, c_fp, c_fp])
    f("denoise", [_VP, C.POINTER(HkDenoiseConfig), c_fp, c_fp])
    if p == "hk_":
        f("read_film_async", [_VP, c_fp, C.POINTER(C.c_int32)])
        f("read_film_wait", [_VP, C.c_int32])
    f("upload_textures", [_VP, C.POINTER(HkTexture), C.c_uint32])
    f("upload_envmaps", [_VP, C.POINTER(HkEnvMap), C.c_uint32])
    f("upload_lights", [_VP, C.POINTER(HkLight), C.c_uint32, C.POINTER(HkLightSampler)])
    f("upload_media", [_VP, C.POINTER(HkMedium), C.c_uint32])
    if p == "hk_":
        f("update_medium", [_VP, C.c_uint32, C.POINTER(HkMedium)])
        f("read_majorant", [_VP, C.c_uint32, c_fp, C.c_uint64])
        f("read_nanovdb", [_VP, C.c_uint32, c_u8p, C.c_uint64, C.POINTER(C.c_uint64)])
    f("set_camera", [_VP, C.POINTER(HkCamera)])
    f("set_filter", [_VP, C.POINTER(HkFilter)])
    f("set_params", [_VP, C.POINTER(HkRenderParams)])
    f("clear", [_VP])
    f("render_samples", [_VP, C.c_int32, C.c_int32])
    f("render_samples_strided", [_VP, C.c_int32, C.c_int32, C.c_int32])
    f("read_film", [_VP, c_fp])
    f("read_accum", [_VP, c_fp, c_fp])
    return lib


def bind_hk(lib):
    bind_common(lib, "hk_")
    def f(name, args, res=C.c_int32):
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    f("hk_abi_version", [])
    f("hk_last_error", [_VP], C.c_char_p)
    f("hk_write_accum", [_VP, c_fp, c_fp])
    f("hk_film_accum_dev", [_VP, C.POINTER(_VP), C.POINTER(C.c_uint64)])
    f("hk_trace_closest", [_VP, c_fp, C.c_uint64, c_fp])
    f("hk_trace_closest_dev", [_VP, _VP, C.c_uint64, _VP, C.c_int32])
    f("hk_trace_any", [_VP, c_fp, C.c_uint64, c_u8p])
    f("hk_stats", [_VP, C.POINTER(HkStats)])
    f("hk_synchronize", [_VP])
    f("hk_read_film_dev", [_VP, _VP])
    f("hk_postprocess_dev", [_VP, C.POINTER(HkPostprocess), _VP])
    f("hk_set_stream", [_VP, _VP])
    f("hk_dev_alloc", [_VP, C.c_uint64, C.POINTER(_VP)])
    f("hk_dev_free", [_VP, _VP])
    f("hk_dev_upload", [_VP, _VP, _VP, C.c_uint64])
    f("hk_dev_download", [_VP, _VP, _VP, C.c_uint64])
    f("hk_pinned_alloc", [C.c_uint64, C.POINTER(_VP)])
    f("hk_pinned_free", [_VP])
    f("hk_set_profiling", [_VP, C.c_int32])
    f("hk_stage_times", [_VP, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
    f("hk_test_trace_counts", [_VP, c_fp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
    # host-side scene-build helpers (CPU code, usable without a GPU)
    f("hk_host_generate_rgb2spec", [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), c_fp, c_fp])
    f("hk_host_build_bvh8", [c_fp, c_u32p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
    f("hk_host_build_light_sampler", [C.POINTER(HkLight), C.c_uint32, C.POINTER(HkLightBVHNode), c_u32p, c_u32p,
                                      c_i32p, c_u32p, c_u32p])
    return lib


_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libhikari_cuda.so")
if os.environ.get("HK_CUDA_LIB"):      # development only: a tuning variant of the same CUDA library (tools/variants.sh)
    LIB_PATH = os.path.abspath(os.environ["HK_CUDA_LIB"])


# The library's frame pipeline runs ~20 streams; the driver's default of 8 hardware work queues makes them serialise (hk_api.cu,
# hk_on_load).  The variable is read when the CUDA context is created, so set it on import, unless the user chose a value.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def load_library():
    """Load libhikari_cuda.so (built in-tree by __graft_entry__.build()).  Fails loudly when missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
        _LIB = bind_hk(C.CDLL(LIB_PATH))
    return _LIB
