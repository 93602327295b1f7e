"""Host-side mirror of the reference's scene / integrator API for the VolPath path.

Julia is not available in this image, so the host side that a Julia user would drive
(`Scene()`, `push!`, `sync!`, `PerspectiveCamera`, `Film`, `VolPath(samples=…, max_depth=…)`,
`integrator(scene, film, camera)`, `render!`, `clear!`) is mirrored here in Python with the same names,
argument meaning and error behaviour, on top of the C ABI in include/hikari_cuda.h.  Everything in this
file is HOST logic (scene flattening, table building); all rendering happens in libhikari_cuda.so.

Reference files mirrored: src/scene.jl, src/scene-mesh.jl, src/materials/*.jl (constructors),
src/lights/*.jl (constructors), src/camera/perspective.jl, src/film.jl:61-183, src/filter.jl:136-150,
611-725, src/sampler/sampling.jl:201-263, src/sampler/sobol.jl:317-323, src/integrators/volpath/volpath.jl:29-113,
445-670.
"""
import ctypes as C
import math
import numpy as np

from . import _abi as A
from . import tables as T

f32 = np.float32


def _fp(a):
    return a.ctypes.data_as(A.c_fp)


def _v3(x):
    return np.asarray(x, dtype=f32).reshape(3)


# ------------------------------------------------------------------------------------------------
# transforms (Raycore.look_at / perspective equivalents; camera looks down -z, see perspective.jl:109)
# ------------------------------------------------------------------------------------------------
def look_at(eye, target, up=(0, 1, 0)):
    """camera_to_world 4x4 (row-major): columns = right, up, back (eye - target), position."""
    eye, target, up = _v3(eye), _v3(target), _v3(up)
    z = eye - target
    z = z / np.linalg.norm(z)
    x = np.cross(up / np.linalg.norm(up), z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4, dtype=f32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return m


def translate(v):
    m = np.eye(4, dtype=f32)
    m[:3, 3] = _v3(v)
    return m


def scale_matrix(s):
    m = np.eye(4, dtype=f32)
    s = np.broadcast_to(np.asarray(s, dtype=f32), (3,))
    m[0, 0], m[1, 1], m[2, 2] = s
    return m


def rotation_matrix(angle_degrees, axis):
    """src/textures/environment_map.jl:52-64 (returns row-major 3x3 of the same matrix)."""
    th = math.radians(angle_degrees)
    a = _v3(axis)
    a = a / np.linalg.norm(a)
    s, c = math.sin(th), math.cos(th)
    t = 1 - c
    return np.array([[t * a[0] * a[0] + c, t * a[0] * a[1] + s * a[2], t * a[0] * a[2] - s * a[1]],
                     [t * a[0] * a[1] - s * a[2], t * a[1] * a[1] + c, t * a[1] * a[2] + s * a[0]],
                     [t * a[0] * a[2] + s * a[1], t * a[1] * a[2] - s * a[0], t * a[2] * a[2] + c]], dtype=f32).T


class Film:
    """src/film.jl:61-183 — only the output-buffer contract used by VolPath."""

    def __init__(self, resolution):
        self.resolution = (int(resolution[0]), int(resolution[1]))   # (W, H)
        w, h = self.resolution
        # The reference's framebuffer is an (H, W) COLUMN-major RGB{Float32} matrix (film.jl:61-106): linear index
        # (px-1)*H + py.  The same bytes are a C-order (W, H, 3) array; `framebuffer` is the [py, px] view of it, so
        # hk_read_film writes straight into it.  The storage is page-locked when a CUDA device is present.
        self._pins = []                                              # (lib, pointer) of every page-locked buffer this film owns
        self._free = []                                              # host buffers neither displayed nor targeted by an un-waited read-out
        self._store = self._new_store()
        self.framebuffer = self._store.transpose(1, 0, 2)            # framebuffer[py, px] = RGB (a view)
        self.iteration_index = 0

    def _new_store(self, zero=True):
        w, h = self.resolution
        store = None
        try:
            lib = A.load_library()
            p = C.c_void_p()
            if lib.hk_pinned_alloc(12 * w * h, C.byref(p)) == 0 and p.value:
                self._pins.append((lib, p))
                store = np.ctypeslib.as_array(C.cast(p, A.c_fp), shape=(w, h, 3))
        except Exception:
            store = None
        if store is None:
            store = np.empty((w, h, 3), dtype=f32)
        if zero:
            store[...] = 0
        return store

    def _acquire_store(self):
        """A host buffer for one pipelined read-out (Backend.read_film_async): never the displayed one and never one that an
        un-waited read-out still targets -- with four frames in flight plus the displayed frame that is five buffers
        (include/hikari_cuda.h: 'the caller alternates host buffers'); they are allocated on demand and recycled by _show."""
        return self._free.pop() if self._free else self._new_store(zero=False)      # (a read-out overwrites every byte)

    def _show(self, store):
        if store is not self._store:
            self._free.append(self._store)
            self._store = store
            self.framebuffer = self._store.transpose(1, 0, 2)

    def __del__(self):
        for lib, p in getattr(self, "_pins", []):
            lib.hk_pinned_free(p)
        self._pins = []
        self.framebuffer = None; self._store = None; self._free = []

    def clear(self):
        self.framebuffer[:] = 0
        self.iteration_index = 0
        for a in getattr(self, "_aux_store", None) or ():       # clear!(film) also resets albedo / normal / depth, film.jl:343-346
            a[...] = 0


class FilmSensor:
    """src/postprocess.jl:37-47: FilmSensor(iso=100, exposure_time=1.0, white_balance=0)"""

    def __init__(self, iso=100.0, exposure_time=1.0, white_balance=0.0):
        self.iso, self.exposure_time, self.white_balance = f32(iso), f32(exposure_time), f32(white_balance)


def planckian_xy(T):
    """src/spectral/color.jl:469-493 (CIE 015:2004), Float32 arithmetic"""
    T = f32(T); T2 = f32(T * T); T3 = f32(T2 * T)
    if T <= f32(4000):
        x = f32(f32(f32(f32(-0.2661239e9) / T3) - f32(f32(0.2343589e6) / T2)) + f32(f32(0.8776956e3) / T)) + f32(0.179910)
    else:
        x = f32(f32(f32(f32(-3.0258469e9) / T3) + f32(f32(2.1070379e6) / T2)) + f32(f32(0.2226347e3) / T)) + f32(0.240390)
    x = f32(x); x2 = f32(x * x); x3 = f32(x2 * x)
    if T <= f32(2222):
        c = (-1.1063814, -1.34811020, 2.18555832, -0.20219683)
    elif T <= f32(4000):
        c = (-0.9549476, -1.37418593, 2.09137015, -0.16748867)
    else:
        c = (3.0817580, -5.87338670, 3.75112997, -0.37001483)
    y = f32(f32(f32(f32(f32(c[0]) * x3) + f32(f32(c[1]) * x2)) + f32(f32(c[2]) * x)) + f32(c[3]))
    return x, y


def compute_white_balance_matrix(src_temp):
    """src/spectral/color.jl:522-546: Bradford adaptation from a Planckian source at src_temp K to D65 (3x3, Float32)"""
    lms_from_xyz = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]], dtype=f32)
    xyz_from_lms = np.array([[0.9869929, -0.1470543, 0.1599627], [0.4323053, 0.5183603, 0.0492912], [-0.0085287, 0.0400428, 0.9684867]], dtype=f32)

    def xy_to_xyz(x, y):
        return np.array([f32(x / y), f32(1.0), f32(f32(f32(f32(1.0) - x) - y) / y)], dtype=f32)

    src = lms_from_xyz @ xy_to_xyz(*planckian_xy(src_temp))
    dst = lms_from_xyz @ xy_to_xyz(f32(0.31272), f32(0.32903))
    scale = np.diag((dst / src).astype(f32)).astype(f32)
    return ((xyz_from_lms @ scale).astype(f32) @ lms_from_xyz).astype(f32)


TONEMAP_MODES = {None: 0, "none": 0, "reinhard": 1, "reinhard_extended": 2, "aces": 3, "uncharted2": 4, "filmic": 5}


def postprocess(film, vp, exposure=1.0, tonemap="aces", gamma=2.2, white_point=4.0, sensor=None, background=None):
    """postprocess!(film; exposure, tonemap, gamma, white_point, sensor), src/postprocess.jl:281-357.  Non-destructive:
    reads the accumulated film of `vp`'s backend, writes film.postprocess ([py, px] like film.framebuffer).  The reference
    reads film.framebuffer; here the division by the weight sum is fused into the same pass (hk_postprocess)."""
    if tonemap not in TONEMAP_MODES:
        raise ValueError(f"unknown tonemap {tonemap!r}")
    sensor = sensor if sensor is not None else FilmSensor()
    p = A.HkPostprocess()
    p.exposure = float(f32(exposure)); p.tonemap_mode = TONEMAP_MODES[tonemap]
    p.apply_gamma = 0 if gamma is None else 1
    p.inv_gamma = 1.0 if gamma is None else float(f32(1.0) / f32(gamma))
    p.white_point = float(f32(white_point))
    p.imaging_ratio = float(f32(f32(sensor.exposure_time * sensor.iso) / f32(100.0)))
    p.apply_wb = 1 if sensor.white_balance > 0 else 0
    wb = compute_white_balance_matrix(sensor.white_balance) if p.apply_wb else np.eye(3, dtype=f32)
    p.wb[:] = [float(v) for v in wb.reshape(-1)]
    p.mask_escaped = 0 if background is None else 1      # escaped pixels (film.depth = Inf) fade to `background`, :339-342
    p.background[:] = (0.0, 0.0, 0.0) if background is None else _rgb(background)
    w, h = film.resolution
    if getattr(film, "_pp_store", None) is None:
        film._pp_store = np.zeros((w, h, 3), dtype=f32)
        film.postprocess = film._pp_store.transpose(1, 0, 2)
    vp.backend.call("postprocess", C.byref(p), _fp(film._pp_store))
    return film.postprocess


def fill_aux_buffers(film, vp, has_infinite_lights=False):
    """fill_aux_buffers!(film, scene, camera; has_infinite_lights), src/film.jl:410-431: film.albedo / film.normal ([py, px, 3])
    and film.depth ([py, px]) from one centre-of-pixel primary ray per pixel, traced by `vp`'s backend against the scene and
    camera it last rendered (or was given through set_camera)."""
    w, h = film.resolution
    if getattr(film, "_aux_store", None) is None:
        film._aux_store = (np.zeros((w, h, 3), dtype=f32), np.zeros((w, h, 3), dtype=f32), np.zeros((w, h), dtype=f32))
        film.albedo, film.normal = film._aux_store[0].transpose(1, 0, 2), film._aux_store[1].transpose(1, 0, 2)
        film.depth = film._aux_store[2].transpose(1, 0)
    vp.backend.call("fill_aux_buffers", 1 if has_infinite_lights else 0)
    vp.backend.call("read_aux_buffers", _fp(film._aux_store[0]), _fp(film._aux_store[1]), _fp(film._aux_store[2]))
    return film


class DenoiseConfig:
    """src/denoise.jl:29-57"""

    def __init__(self, iterations=5, sigma_color=4.0, sigma_normal=128.0, sigma_depth=1.0, use_variance=True):
        self.iterations, self.use_variance = int(iterations), bool(use_variance)
        self.sigma_color, self.sigma_normal, self.sigma_depth = float(sigma_color), float(sigma_normal), float(sigma_depth)


def denoise(film, vp, config=None):
    """denoise!(film; config), src/denoise.jl:301-372: edge-avoiding a-trous filter of the framebuffer guided by film.normal /
    film.depth (fill_aux_buffers first); result in film.postprocess.  As in the reference, from two iterations on
    film.framebuffer is left holding the last even pass (it is one of the two ping-pong buffers)."""
    config = config if config is not None else DenoiseConfig()
    c = A.HkDenoiseConfig(config.iterations, config.sigma_color, config.sigma_normal, config.sigma_depth, 1 if config.use_variance else 0)
    w, h = film.resolution
    if getattr(film, "_pp_store", None) is None:
        film._pp_store = np.zeros((w, h, 3), dtype=f32)
        film.postprocess = film._pp_store.transpose(1, 0, 2)
    vp.backend.call("denoise", C.byref(c), _fp(film._pp_store), _fp(film._store))
    return None


def denoise_inplace(film, vp, config=None):
    """denoise_inplace!(film; config), src/denoise.jl:379-383"""
    denoise(film, vp, config)
    film.framebuffer[:] = film.postprocess
    return None


class PerspectiveCamera:
    """src/camera/perspective.jl:41-91.  `screen_window=None` is the reference convenience constructor's literal
    [-1,1]^2 window (:84: non-square films are stretched, exactly as upstream); "aspect" picks the aspect-correct window
    (shorter axis spans [-1,1]) that the full constructor is normally called with (examples/*.jl, RayMakie)."""

    def __init__(self, eyepos, lookat, film, up=(0, 1, 0), fov=55.0, lens_radius=0.0, focal_distance=1e6,
                 screen_window=None):
        w, h = film.resolution
        self.camera_to_world = look_at(eyepos, lookat, up)
        if screen_window is None:
            screen_window = ((-1.0, -1.0), (1.0, 1.0))
        elif isinstance(screen_window, str):
            assert screen_window == "aspect"
            aspect = w / h
            screen_window = ((-aspect, -1.0), (aspect, 1.0)) if aspect >= 1 else ((-1.0, -1 / aspect), (1.0, 1 / aspect))
        (x0, y0), (x1, y1) = screen_window
        near = 0.01
        tan_half = math.tan(math.radians(fov) / 2)
        sx, sy = (x1 - x0) / w, (y1 - y0) / h
        m = np.zeros((4, 4), dtype=np.float64)
        m[0, 0], m[0, 3] = sx * tan_half * near, x0 * tan_half * near
        m[1, 1], m[1, 3] = sy * tan_half * near, y0 * tan_half * near
        m[2, 3] = -near
        m[3, 3] = 1
        self.raster_to_camera = m.astype(f32)
        self.lens_radius, self.focal_distance = float(lens_radius), float(focal_distance)
        self.shutter_open, self.shutter_close = 0.0, 1.0
        self.dx_camera = np.array([m[0, 0], 0, 0], dtype=f32)
        self.dy_camera = np.array([0, m[1, 1], 0], dtype=f32)
        self.position = _v3(eyepos)

    def to_abi(self):
        c = A.HkCamera()
        c.raster_to_camera[:] = self.raster_to_camera.reshape(-1).tolist()
        c.camera_to_world[:] = self.camera_to_world.reshape(-1).tolist()
        c.lens_radius, c.focal_distance = self.lens_radius, self.focal_distance
        c.shutter_open, c.shutter_close = self.shutter_open, self.shutter_close
        c.dx_camera[:] = self.dx_camera.tolist()
        c.dy_camera[:] = self.dy_camera.tolist()
        return c


# ------------------------------------------------------------------------------------------------
# filters (src/filter.jl)
# ------------------------------------------------------------------------------------------------
class GaussianFilter:
    """filter.jl:136-162"""

    def __init__(self, radius=(1.5, 1.5), sigma=0.5):
        self.radius = (f32(radius[0]), f32(radius[1]))
        self.sigma = f32(sigma)
        self.exp_x = self._g(self.radius[0])
        self.exp_y = self._g(self.radius[1])
        self.type = 3

    def _g(self, x):
        x = f32(x)
        return f32(np.exp(f32(-(x * x)) / f32(f32(2) * self.sigma * self.sigma)))

    def evaluate(self, px, py):
        gx = max(f32(0), f32(self._g(px) - self.exp_x))
        gy = max(f32(0), f32(self._g(py) - self.exp_y))
        return f32(gx * gy)


class BoxFilter:
    def __init__(self, radius=(0.5, 0.5)):
        self.radius = (f32(radius[0]), f32(radius[1]))
        self.type = 1


class TriangleFilter:
    def __init__(self, radius=(2.0, 2.0)):
        self.radius = (f32(radius[0]), f32(radius[1]))
        self.type = 2


class FilterSamplerData:
    """GPUFilterSamplerData, filter.jl:611-725 (all arithmetic in Float32, sequential sums)."""

    def __init__(self, flt):
        self.flt = flt
        self.keep = []
        if flt.type in (1, 2):
            self.nx = self.ny = 0
            return
        r = flt.radius
        nx = max(int(math.ceil(32 * float(r[0]))), 8)
        ny = max(int(math.ceil(32 * float(r[1]))), 8)
        dmin = (f32(-r[0]), f32(-r[1]))
        dmax = (f32(r[0]), f32(r[1]))
        dx = f32(f32(dmax[0] - dmin[0]) / f32(nx))
        dy = f32(f32(dmax[1] - dmin[1]) / f32(ny))
        func = np.zeros((ny, nx), dtype=f32)
        for iy in range(1, ny + 1):
            py = f32(dmin[1] + f32(f32(f32(iy) - f32(0.5)) * dy))
            for ix in range(1, nx + 1):
                px = f32(dmin[0] + f32(f32(f32(ix) - f32(0.5)) * dx))
                func[iy - 1, ix - 1] = max(f32(0), flt.evaluate(px, py))
        marginal_func = np.zeros(ny, dtype=f32)
        for iy in range(ny):
            acc = f32(0)
            for ix in range(nx):
                acc = f32(acc + func[iy, ix])
            marginal_func[iy] = acc
        mcdf = np.zeros(ny + 1, dtype=f32)
        for iy in range(ny):
            mcdf[iy + 1] = f32(mcdf[iy] + marginal_func[iy])
        self.func_integral = f32(f32(mcdf[-1] * dx) * dy)
        end = mcdf[-1]
        if end > 0:
            mcdf = (mcdf / end).astype(f32)
        else:
            mcdf = (np.arange(ny + 1, dtype=f32) / f32(ny)).astype(f32)
        ccdf = np.zeros((ny, nx + 1), dtype=f32)
        for iy in range(ny):
            for ix in range(nx):
                ccdf[iy, ix + 1] = f32(ccdf[iy, ix] + func[iy, ix])
            rs = ccdf[iy, nx]
            if rs > 0:
                ccdf[iy, :] = (ccdf[iy, :] / rs).astype(f32)
            else:
                ccdf[iy, :] = (np.arange(nx + 1, dtype=f32) / f32(nx)).astype(f32)
        self.nx, self.ny = nx, ny
        self.func, self.marginal_cdf, self.marginal_func, self.conditional_cdf = func, mcdf, marginal_func, ccdf
        self.domain_min, self.domain_max = dmin, dmax

    def to_abi(self):
        f = A.HkFilter()
        f.type = self.flt.type
        f.radius[:] = [float(self.flt.radius[0]), float(self.flt.radius[1])]
        f.nx, f.ny = self.nx, self.ny
        if self.nx:
            f.func, f.marginal_cdf = _fp(self.func), _fp(self.marginal_cdf)
            f.marginal_func, f.conditional_cdf = _fp(self.marginal_func), _fp(self.conditional_cdf)
            f.domain_min[:] = [float(self.domain_min[0]), float(self.domain_min[1])]
            f.domain_max[:] = [float(self.domain_max[0]), float(self.domain_max[1])]
            f.func_integral = float(self.func_integral)
        return f


# ------------------------------------------------------------------------------------------------
# materials (constructors mirror src/materials/*.jl keyword constructors; constant parameters only)
# ------------------------------------------------------------------------------------------------
def _rgb(x):
    if np.isscalar(x):
        return (float(x),) * 3
    x = tuple(float(v) for v in x)
    assert len(x) == 3
    return x


class Material:
    type = 0

    def to_abi(self, scene):
        raise NotImplementedError


class MixMaterial(Material):
    """src/materials/mix-material.jl:39-99: MixMaterial(materials=(m1, m2), amount=0.5).  Resolved to one of the two at
    intersection time by a hash of (hit point, wo, SetKeys); `amount` is a constant or a Texture{Float32} evaluated at the hit's uv
    (choose_material, mix-material.jl:178-196: `amt = eval_tex(ctx, mix.amount, uv)`).
    The reference takes the sub-materials' SetKeys explicitly (`material_indices`); the mirror derives them from the
    scene: type_idx = position of the material's type in first-push order, vec_idx = position within that type."""
    type = A.HK_MAT_MIX

    def __init__(self, materials, amount=0.5):
        assert len(materials) == 2 and all(isinstance(m, Material) for m in materials)
        self.materials, self.amount = tuple(materials), _param_f(amount)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        _put_f(m, scene, 0, self.amount)
        keys = []
        for k, sub in enumerate(self.materials):
            m.ival[k] = scene._index_of(scene.materials, sub)
            keys.append(scene.set_key(sub))
        m.spec[0], m.spec[1] = keys[0][1], keys[1][1]
        m.flags = keys[0][0] | (keys[1][0] << 8)
        return m


Mix = MixMaterial


class Texture:
    """src/textures/basic.jl: Texture(data) over an (h, w) image of RGB values — data[y, x] as in the reference, row 1 at v = 1.
    Stored column-major like the Julia matrix, so the bytes handed to hk_upload_textures are the reference's own."""

    def __init__(self, data_hw3):
        d = np.asarray(data_hw3, dtype=f32)
        if d.ndim == 2:                                                  # Texture{Float32}: uploaded as r = g = b (HkMaterial.ftex reads the first channel)
            d = np.repeat(d[..., None], 3, axis=2)
        assert d.ndim == 3 and d.shape[2] in (3, 4), "RGB(A) texture: [h, w, 3] or [h, w, 4]; scalar texture: [h, w]"
        self.alpha = None        # the reference's RGBSpectrum texels carry alpha as their 4th float (spectrum.jl:62-70)
        if d.shape[2] == 4:
            if not (d[..., 3] == 1).all():
                self.alpha = np.ascontiguousarray(d[..., 3].T)        # column-major (h, w), like the colours
            d = d[..., :3]
        self.h, self.w = int(d.shape[0]), int(d.shape[1])
        self.data = np.ascontiguousarray(d.transpose(1, 0, 2))       # [x][y][3] = column-major (h, w)


class VertexColorTexture(Texture):
    """src/textures/basic.jl:43-46: three corner colours per face, face_colors[n_faces, 3, 3] (face, corner, rgb); evaluated by
    barycentric interpolation with the hit's face index (texture-ref.jl:240-245).  Stored as the reference's (3, n_faces) matrix."""

    def __init__(self, face_colors):
        d = np.asarray(face_colors, dtype=f32)
        assert d.ndim == 3 and d.shape[1:] == (3, 3), "face_colors: [n_faces, 3 corners, 3]"
        self.h, self.w, self.alpha = 3, int(d.shape[0]), None
        self.data = np.ascontiguousarray(d)                          # [face][corner][3] = column-major (3, n_faces)


def _param_rgb(v):
    """an RGB material parameter: a constant or a Texture (the reference's fields are `Texture, Raycore.TextureRef, or raw RGBSpectrum`)"""
    return v if isinstance(v, Texture) else _rgb(v)


def _put_rgb(m, scene, slot, v):
    """HkMaterial.rgb<slot> = the constant, or HkMaterial.tex[slot] = the texture's id (eval_tex at the hit, texture-ref.jl:72-80)"""
    if isinstance(v, Texture):
        assert not isinstance(v, VertexColorTexture), "VertexColorTexture is supported for MatteMaterial.Kd only"
        m.tex[slot] = scene._texture_id(v)
    else:
        (m.rgb0, m.rgb1, m.rgb2)[slot][0:3] = _rgb(v)


def _param_f(v):
    return v if isinstance(v, Texture) else float(v)


def _put_f(m, scene, k, v):
    if isinstance(v, Texture):
        m.ftex[k] = scene._texture_id(v)
    else:
        m.f[k] = float(v)


class MatteMaterial(Material):            # uber-material.jl:180-183, 256
    type = A.HK_MAT_MATTE

    def __init__(self, Kd=0.5, sigma=0.0):
        self.Kd, self.sigma = (Kd if isinstance(Kd, Texture) else _rgb(Kd)), _param_f(sigma)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        if isinstance(self.Kd, Texture):
            m.tex[0] = scene._texture_id(self.Kd)                     # a TextureRef in the reference (texture-ref.jl:196-199)
            if isinstance(self.Kd, VertexColorTexture):
                m.flags |= A.HK_MATFLAG_VERTEX_COLORS
        else:
            m.rgb0[:] = self.Kd
        _put_f(m, scene, 0, self.sigma)
        return m


class MirrorMaterial(Material):           # :193-195, 275
    type = A.HK_MAT_MIRROR

    def __init__(self, Kr=0.9):
        self.Kr = _param_rgb(Kr)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        _put_rgb(m, scene, 0, self.Kr)
        return m


class GlassMaterial(Material):            # :209-216, 299
    type = A.HK_MAT_GLASS

    def __init__(self, Kr=1.0, Kt=1.0, index=1.5, u_roughness=0.0, v_roughness=0.0, remap_roughness=True):
        self.Kr, self.Kt, self.index = _param_rgb(Kr), _param_rgb(Kt), _param_f(index)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        _put_rgb(m, scene, 0, self.Kr)
        _put_rgb(m, scene, 1, self.Kt)
        _put_f(m, scene, 0, self.index)
        return m


class PiecewiseLinearSpectrum:            # src/spectral/piecewise-linear.jl:4-7
    def __init__(self, lambdas, values):
        self.lambdas = np.asarray(lambdas, dtype=f32)
        self.values = np.asarray(values, dtype=f32)


class ConductorMaterial(Material):        # :378-384, 418-426
    type = A.HK_MAT_CONDUCTOR

    def __init__(self, eta=(0.2, 0.2, 0.2), k=(3.9, 3.9, 3.9), roughness=0.1, reflectance=(1, 1, 1),
                 remap_roughness=True):
        self.eta, self.k = eta, k
        self.roughness, self.reflectance, self.remap = _param_f(roughness), _rgb(reflectance), bool(remap_roughness)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        m.flags = A.HK_MATFLAG_REMAP_ROUGHNESS if self.remap else 0
        _put_f(m, scene, 0, self.roughness)
        spectral = isinstance(self.eta, PiecewiseLinearSpectrum)
        assert spectral == isinstance(self.k, PiecewiseLinearSpectrum), "eta and k must both be spectra or both RGB"
        if spectral:
            m.flags |= A.HK_MATFLAG_SPECTRAL_ETA_K
            m.spec[0] = scene._spectrum_id(self.eta)
            m.spec[1] = scene._spectrum_id(self.k)
        else:
            _put_rgb(m, scene, 0, self.eta)
            _put_rgb(m, scene, 1, self.k)
        return m


def _metal(name, roughness, remap):
    t = T.load_tables()
    e, k = t[name + "_eta"], t[name + "_k"]
    return ConductorMaterial(PiecewiseLinearSpectrum(e[0], e[1]), PiecewiseLinearSpectrum(k[0], k[1]), roughness,
                             (1, 1, 1), remap)


def Gold(roughness=0.0, remap_roughness=True):      # uber-material.jl:469-470
    return _metal("au", roughness, remap_roughness)


def Silver(roughness=0.0, remap_roughness=True):
    return _metal("ag", roughness, remap_roughness)


def Copper(roughness=0.0, remap_roughness=True):
    return _metal("cu", roughness, remap_roughness)


def Aluminum(roughness=0.0, remap_roughness=True):
    return _metal("al", roughness, remap_roughness)


class CoatedDiffuseMaterial(Material):    # coated-diffuse.jl:98-127
    type = A.HK_MAT_COATED_DIFFUSE

    def __init__(self, reflectance=0.5, roughness=0.0, thickness=0.01, eta=1.5, albedo=0.0, g=0.0, max_depth=10,
                 n_samples=1, remap_roughness=True):
        self.reflectance, self.albedo = _param_rgb(reflectance), _param_rgb(albedo)
        self.u_rough, self.v_rough = (roughness if isinstance(roughness, tuple) else (roughness, roughness))
        self.thickness, self.eta, self.g = _param_f(thickness), float(eta), _param_f(g)
        self.max_depth, self.n_samples, self.remap = int(max_depth), int(n_samples), bool(remap_roughness)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        m.flags = A.HK_MATFLAG_REMAP_ROUGHNESS if self.remap else 0
        _put_rgb(m, scene, 0, self.reflectance)
        _put_rgb(m, scene, 1, self.albedo)
        for k, v in enumerate((self.u_rough, self.v_rough, self.thickness, self.eta, self.g)):
            _put_f(m, scene, k, v)
        m.ival[0], m.ival[1] = self.max_depth, self.n_samples
        return m


class CoatedDiffuseTransmissionMaterial(CoatedDiffuseMaterial):   # coated-diffuse-transmission.jl (keyword constructor)
    """Dielectric coating over a Lambertian base that reflects (`reflectance`) and transmits (`transmittance`)."""
    type = A.HK_MAT_COATED_DIFFUSE_TRANSMISSION

    def __init__(self, reflectance=0.5, transmittance=0.25, roughness=0.0, thickness=0.01, eta=1.5, albedo=0.0, g=0.0,
                 max_depth=10, n_samples=1, remap_roughness=True):
        super().__init__(reflectance, roughness, thickness, eta, albedo, g, max_depth, n_samples, remap_roughness)
        self.transmittance = _param_rgb(transmittance)

    def to_abi(self, scene):
        m = super().to_abi(scene)
        _put_rgb(m, scene, 2, self.transmittance)
        return m


CoatedDiffuseTransmission = CoatedDiffuseTransmissionMaterial


class CoatedConductorMaterial(Material):  # coated-conductor.jl:48-105 (struct), 167-243 (keyword constructor)
    """Dielectric coating over a conductor.  Give `conductor_eta` + `conductor_k` (RGB or PiecewiseLinearSpectrum) or the
    artist `reflectance`; with neither, reflectance = 1 as in the reference."""
    type = A.HK_MAT_COATED_CONDUCTOR

    def __init__(self, interface_roughness=0.0, interface_eta=1.5, conductor_eta=None, conductor_k=None, reflectance=None,
                 conductor_roughness=0.01, thickness=0.01, albedo=0.0, g=0.0, max_depth=10, n_samples=1,
                 remap_roughness=True):
        if conductor_eta is not None and conductor_k is None:
            raise ValueError("conductor_k must be provided when using conductor_eta")     # coated-conductor.jl:203-205
        pair = lambda r: tuple(_param_f(v) for v in r) if isinstance(r, tuple) else (_param_f(r), _param_f(r))
        self.i_rough, self.c_rough = pair(interface_roughness), pair(conductor_roughness)
        self.interface_eta, self.thickness, self.g = float(interface_eta), _param_f(thickness), _param_f(g)
        self.use_eta_k = conductor_eta is not None
        self.conductor_eta, self.conductor_k = conductor_eta, conductor_k
        self.reflectance = _param_rgb(1.0 if reflectance is None else reflectance)
        self.albedo = _param_rgb(albedo)
        self.max_depth, self.n_samples, self.remap = int(max_depth), int(n_samples), bool(remap_roughness)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        m.flags = A.HK_MATFLAG_REMAP_ROUGHNESS if self.remap else 0
        if self.use_eta_k:
            m.flags |= A.HK_MATFLAG_USE_ETA_K
            spectral = isinstance(self.conductor_eta, PiecewiseLinearSpectrum)
            assert spectral == isinstance(self.conductor_k, PiecewiseLinearSpectrum), "eta and k must both be spectra or both RGB"
            if spectral:
                m.flags |= A.HK_MATFLAG_SPECTRAL_ETA_K
                m.spec[0], m.spec[1] = scene._spectrum_id(self.conductor_eta), scene._spectrum_id(self.conductor_k)
            else:
                _put_rgb(m, scene, 0, self.conductor_eta)
                _put_rgb(m, scene, 1, self.conductor_k)
        else:
            _put_rgb(m, scene, 0, self.reflectance)
        _put_rgb(m, scene, 2, self.albedo)
        for k, v in enumerate((self.i_rough[0], self.i_rough[1], self.thickness, self.interface_eta, self.g, self.c_rough[0], self.c_rough[1])):
            _put_f(m, scene, k, v)
        m.ival[0], m.ival[1] = self.max_depth, self.n_samples
        return m


CoatedConductor = CoatedConductorMaterial      # coated-conductor.jl:249


class ThinDielectricMaterial(Material):   # thin-dielectric.jl:45-52
    type = A.HK_MAT_THIN_DIELECTRIC

    def __init__(self, eta=1.5):
        self.eta = float(eta)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        m.f[0] = self.eta
        return m


class DiffuseTransmissionMaterial(Material):   # diffuse-transmission.jl:39-85
    type = A.HK_MAT_DIFFUSE_TRANSMISSION

    def __init__(self, reflectance=0.25, transmittance=0.25, scale=1.0):
        self.reflectance, self.transmittance, self.scale = _param_rgb(reflectance), _param_rgb(transmittance), float(scale)

    def to_abi(self, scene):
        m = A.HkMaterial(type=self.type)
        _put_rgb(m, scene, 0, self.reflectance)
        _put_rgb(m, scene, 1, self.transmittance)
        m.f[0] = self.scale
        return m


class MediumInterface:
    """src/materials/medium-interface.jl:40-56.  `emission=(Le_rgb, scale, two_sided)` registers one
    DiffuseAreaLight per face (src/scene-mesh.jl:100-138)."""

    def __init__(self, material, inside=None, outside=None, emission=None):
        self.material, self.inside, self.outside, self.emission = material, inside, outside, emission


# ------------------------------------------------------------------------------------------------
# media (src/integrators/volpath/media.jl, nanovdb.jl)
# ------------------------------------------------------------------------------------------------
class HomogeneousMedium:                  # media.jl:735-750
    def __init__(self, sigma_a=0.01, sigma_s=1.0, Le=0.0, g=0.0):
        self.sigma_a, self.sigma_s, self.Le, self.g = _rgb(sigma_a), _rgb(sigma_s), _rgb(Le), float(g)

    def to_abi(self, keep, device_majorant=False):
        m = A.HkMedium(type=A.HK_MEDIUM_HOMOGENEOUS)
        m.sigma_a_rgb[:], m.sigma_s_rgb[:], m.Le_rgb[:] = self.sigma_a, self.sigma_s, self.Le
        m.g, m.scale = self.g, 1.0
        return m


def build_majorant_grid(density_xyz, res):
    """media.jl:1458-1487; density indexed [x, y, z]; returns array [rz][ry][rx]."""
    nx, ny, nz = density_xyz.shape
    out = np.zeros((res[2], res[1], res[0]), dtype=f32)
    def rng(i, n, r):
        s = max(1, int(math.floor(i * n / r)) + 1)
        e = min(n, int(math.ceil((i + 1) * n / r)))
        return s - 1, e
    for iz in range(res[2]):
        z0, z1 = rng(iz, nz, res[2])
        for iy in range(res[1]):
            y0, y1 = rng(iy, ny, res[1])
            for ix in range(res[0]):
                x0, x1 = rng(ix, nx, res[0])
                blk = density_xyz[x0:x1, y0:y1, z0:z1]
                out[iz, iy, ix] = max(f32(0), blk.max()) if blk.size else f32(0)
    return out


class GridMedium:                         # media.jl:886-936
    def __init__(self, density_xyz, sigma_a=0.01, sigma_s=1.0, g=0.0, bounds=((0, 0, 0), (1, 1, 1)),
                 transform=None, majorant_res=(16, 16, 16)):
        self.density = np.ascontiguousarray(np.asarray(density_xyz, dtype=f32))
        self.sigma_a, self.sigma_s, self.g = _rgb(sigma_a), _rgb(sigma_s), float(g)
        self.bounds = (np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32))
        self.medium_to_render = np.eye(4, dtype=f32) if transform is None else np.asarray(transform, dtype=f32)
        self.render_to_medium = np.linalg.inv(self.medium_to_render.astype(np.float64)).astype(f32)
        self.majorant_res = tuple(int(v) for v in majorant_res)
        self._majorant = None

    @property
    def majorant(self):
        """build_majorant_grid on the host (numpy): what the oracle is handed, and what the device-built grid is tested against;
        the CUDA back end builds its own from the uploaded voxels (HkMedium.majorant = NULL)"""
        if self._majorant is None:
            self._majorant = build_majorant_grid(self.density, self.majorant_res)
        return self._majorant

    def to_abi(self, keep, device_majorant=False):
        m = A.HkMedium(type=A.HK_MEDIUM_GRID)
        m.sigma_a_rgb[:], m.sigma_s_rgb[:], m.Le_rgb[:] = self.sigma_a, self.sigma_s, (0, 0, 0)
        m.g, m.scale = self.g, 1.0
        m.bounds_min[:], m.bounds_max[:] = self.bounds[0].tolist(), self.bounds[1].tolist()
        m.render_from_medium[:] = self.medium_to_render.reshape(-1).tolist()
        m.medium_from_render[:] = self.render_to_medium.reshape(-1).tolist()
        nx, ny, nz = self.density.shape
        m.density_res[:] = [nx, ny, nz]
        d = np.ascontiguousarray(self.density.transpose(2, 1, 0))     # -> [nz][ny][nx]
        keep.append(d)
        m.density = _fp(d)
        m.majorant_res[:] = list(self.majorant_res)
        if not device_majorant:
            mj = np.ascontiguousarray(self.majorant)
            keep.append(mj)
            m.majorant = _fp(mj)
        return m


class RGBGridMedium:                      # media.jl:1002-1114
    """Per-voxel RGB absorption / scattering (and optional emission) coefficients, pbrt-v4's RGBGridMedium.  Grids are
    [nx, ny, nz, 3] arrays (the reference's Array{RGBSpectrum,3}); at least one of sigma_a_grid / sigma_s_grid is required,
    an absent one counts as 1 everywhere; Le_grid requires sigma_a_grid."""

    def __init__(self, sigma_a_grid=None, sigma_s_grid=None, Le_grid=None, sigma_scale=1.0, Le_scale=0.0, g=0.0,
                 bounds=((0, 0, 0), (1, 1, 1)), transform=None, majorant_res=(16, 16, 16)):
        conv = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=f32))
        self.sigma_a_grid, self.sigma_s_grid, self.Le_grid = conv(sigma_a_grid), conv(sigma_s_grid), conv(Le_grid)
        if self.sigma_a_grid is None and self.sigma_s_grid is None:
            raise ValueError("At least one of sigma_a_grid or sigma_s_grid must be provided")            # media.jl:1073
        if self.Le_grid is not None and self.sigma_a_grid is None:
            raise ValueError("Le_grid requires sigma_a_grid to be provided (following pbrt-v4)")          # :1075
        shapes = {a.shape for a in (self.sigma_a_grid, self.sigma_s_grid, self.Le_grid) if a is not None}
        if len(shapes) != 1 or len(next(iter(shapes))) != 4 or next(iter(shapes))[3] != 3:
            raise ValueError("RGB grids must share one [nx, ny, nz, 3] shape")                            # :1078-1083
        self.grid_res = next(iter(shapes))[:3]
        self.sigma_scale, self.Le_scale, self.g = float(sigma_scale), float(Le_scale), float(g)
        self.bounds = (np.asarray(bounds[0], dtype=f32), np.asarray(bounds[1], dtype=f32))
        self.medium_to_render = np.eye(4, dtype=f32) if transform is None else np.asarray(transform, dtype=f32)
        self.render_to_medium = np.linalg.inv(self.medium_to_render.astype(np.float64)).astype(f32)
        self.majorant_res = tuple(int(v) for v in majorant_res)
        self._majorant = None

    @property
    def majorant(self):
        """build_rgb_majorant_grid, media.jl:1122-1183, on the host: sigma_scale * (max sigma_a + max sigma_s) per coarse voxel, the
        max taken over the voxel block AND the three channels; an absent grid contributes 1"""
        if self._majorant is None:
            one = np.ones(self.grid_res, dtype=f32)
            ma = build_majorant_grid(self.sigma_a_grid.max(axis=3), self.majorant_res) if self.sigma_a_grid is not None else build_majorant_grid(one, self.majorant_res)
            ms = build_majorant_grid(self.sigma_s_grid.max(axis=3), self.majorant_res) if self.sigma_s_grid is not None else build_majorant_grid(one, self.majorant_res)
            self._majorant = (f32(self.sigma_scale) * (ma + ms)).astype(f32)
        return self._majorant

    def to_abi(self, keep, device_majorant=False):
        m = A.HkMedium(type=A.HK_MEDIUM_RGBGRID)
        m.g, m.scale, m.Le_scale = self.g, self.sigma_scale, self.Le_scale
        m.bounds_min[:], m.bounds_max[:] = self.bounds[0].tolist(), self.bounds[1].tolist()
        m.render_from_medium[:] = self.medium_to_render.reshape(-1).tolist()
        m.medium_from_render[:] = self.render_to_medium.reshape(-1).tolist()
        m.density_res[:] = list(self.grid_res)
        for name, grid in (("rgb_sigma_a", self.sigma_a_grid), ("rgb_sigma_s", self.sigma_s_grid), ("rgb_Le", self.Le_grid)):
            if grid is not None:
                d = np.ascontiguousarray(grid.transpose(2, 1, 0, 3))       # -> [nz][ny][nx][3]
                keep.append(d)
                setattr(m, name, _fp(d))
        m.majorant_res[:] = list(self.majorant_res)
        if not device_majorant:
            mj = np.ascontiguousarray(self.majorant)
            keep.append(mj)
            m.majorant = _fp(mj)
        return m


from .nanovdb import NanoVDBMedium  # noqa: E402  (host-side NanoVDB builder, nanovdb.jl:602-858)


# ------------------------------------------------------------------------------------------------
# lights (constructors mirror src/lights/*.jl)
# ------------------------------------------------------------------------------------------------
class _Light:
    type = 0
    infinite = False

    def _spectrum(self, L, rgb, legacy_rgbspectrum):
        """RGB{Float32} ctor -> RGBIlluminantSpectrum + scale 1/D65_PHOTOMETRIC (point.jl:56-66);
        RGBSpectrum ctor (legacy) keeps the RGB and uplifts on device (point.jl:72-75)."""
        if legacy_rgbspectrum:
            L.spectrum_kind = A.HK_SPECTRUM_RGB
            L.rgb[:] = _rgb(rgb)
        else:
            poly, s = T.rgb_illuminant_spectrum(_rgb(rgb))
            L.spectrum_kind = A.HK_SPECTRUM_ILLUMINANT
            L.poly[:] = [float(v) for v in poly]
            L.illum_scale = float(s)
            L.rgb[:] = _rgb(rgb)
        L.scale = float(f32(1.0) / f32(10567.0))


class PointLight(_Light):
    type = A.HK_LIGHT_POINT

    def __init__(self, rgb, position, legacy_rgbspectrum=False, scale=None):
        """PointLight(rgb::RGB, position) [point.jl:56-66]; legacy_rgbspectrum=True mirrors
        PointLight(i::RGBSpectrum, position) [:72-75, scale 1/D65_PHOTOMETRIC] or, with scale given,
        PointLight(position, i::S, scale) [:26-28]."""
        self.rgb, self.position, self.legacy, self.scale = _rgb(rgb), _v3(position), legacy_rgbspectrum, scale

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        self._spectrum(L, self.rgb, self.legacy)
        if self.scale is not None:
            L.scale = float(self.scale)
        L.position[:] = self.position.tolist()
        return L


class SpotLight(_Light):
    """src/lights/spot.jl.  SpotLight(rgb, position, target, total_width, falloff_start; power=nothing) [:78-94: RGB ->
    RGBIlluminantSpectrum, scale = 1 / spectrum_to_photometric, optional radiant power] or, with legacy_rgbspectrum=True,
    SpotLight(position, target, i::RGBSpectrum, total_width, falloff_start, scale=1) [:49-56].  Angles in degrees; the cone
    points along +z of the light's frame (_spotlight_transform, :104-121)."""
    type = A.HK_LIGHT_SPOT

    def __init__(self, rgb, position, target, total_width, falloff_start, power=None, legacy_rgbspectrum=False, scale=None):
        self.rgb, self.position, self.target = _rgb(rgb), _v3(position), _v3(target)
        self.total_width, self.falloff_start = f32(total_width), f32(falloff_start)
        self.power, self.legacy, self.scale = power, legacy_rgbspectrum, scale
        self.cos_total_width = f32(np.cos(np.deg2rad(self.total_width, dtype=f32), dtype=f32))
        self.cos_falloff_start = f32(np.cos(np.deg2rad(self.falloff_start, dtype=f32), dtype=f32))
        d = (self.target - self.position).astype(f32)
        d = (d * (f32(1) / f32(np.sqrt(f32(np.dot(d, d)))))).astype(f32)
        up = np.array([0, 1, 0], dtype=f32) if abs(d[1]) < f32(0.99) else np.array([1, 0, 0], dtype=f32)
        x = np.cross(up, d).astype(f32)
        x = (x * (f32(1) / f32(np.sqrt(f32(np.dot(x, x)))))).astype(f32)
        y = np.cross(d, x).astype(f32)
        rot = np.eye(4, dtype=f32)
        rot[:3, 0], rot[:3, 1], rot[:3, 2] = x, y, d                # columns = where the local axes land in world space
        w2l = np.linalg.inv(rot).astype(f32) @ translate(-self.position)
        self.world_to_light = w2l.astype(f32)

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        self._spectrum(L, self.rgb, self.legacy)
        if self.legacy:
            L.scale = 1.0 if self.scale is None else float(self.scale)       # spot.jl:49-52: scale defaults to 1
        elif self.power is not None:                                          # spot.jl:86-91
            k_e = f32(2) * f32(np.pi) * ((f32(1) - self.cos_falloff_start) + (self.cos_falloff_start - self.cos_total_width) / f32(2))
            L.scale = float(f32(L.scale) * (f32(self.power) / k_e))
        L.position[:] = self.position.tolist()
        L.cos_total_width, L.cos_falloff_start = float(self.cos_total_width), float(self.cos_falloff_start)
        L.world_to_light[:] = self.world_to_light.reshape(-1).tolist()
        return L


class DirectionalLight(_Light):
    type = A.HK_LIGHT_DIRECTIONAL
    infinite = True

    def __init__(self, rgb, direction, legacy_rgbspectrum=False):
        d = _v3(direction).astype(np.float64)
        self.rgb, self.direction, self.legacy = _rgb(rgb), (d / np.linalg.norm(d)).astype(f32), legacy_rgbspectrum

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        self._spectrum(L, self.rgb, self.legacy)
        L.direction[:] = self.direction.tolist()
        return L


class SunLight(DirectionalLight):
    type = A.HK_LIGHT_SUN


class AmbientLight(_Light):
    type = A.HK_LIGHT_AMBIENT
    infinite = True

    def __init__(self, rgb, legacy_rgbspectrum=False):
        self.rgb, self.legacy = _rgb(rgb), legacy_rgbspectrum

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        self._spectrum(L, self.rgb, self.legacy)
        if self.legacy:
            L.scale = 1.0      # AmbientLight(s::Spectrum) = AmbientLight(s, 1f0), ambient.jl:33
        return L


class Distribution2D:
    """src/sampler/sampling.jl:201-263 (Float32, sequential sums); func has shape (nv, nu)."""

    def __init__(self, func):
        func = np.asarray(func, dtype=f32)
        nv, nu = func.shape
        self.nu, self.nv = nu, nv
        self.conditional_func = np.ascontiguousarray(func)
        ccdf = np.zeros((nv, nu + 1), dtype=f32)
        cfi = np.zeros(nv, dtype=f32)
        for v in range(nv):
            ccdf[v, 1:] = np.cumsum((func[v] / f32(nu)).astype(f32), dtype=f32)
            fi = ccdf[v, nu]
            cfi[v] = fi
            if fi == 0:
                ccdf[v, 1:] = (np.arange(1, nu + 1, dtype=f32) / f32(nu)).astype(f32)
            else:
                ccdf[v, 1:] = (ccdf[v, 1:] / fi).astype(f32)
        self.conditional_cdf, self.conditional_func_int = ccdf, cfi
        self.marginal_func = cfi.copy()
        mcdf = np.zeros(nv + 1, dtype=f32)
        mcdf[1:] = np.cumsum((cfi / f32(nv)).astype(f32), dtype=f32)
        self.marginal_func_int = f32(mcdf[nv])
        if self.marginal_func_int == 0:
            mcdf[1:] = (np.arange(1, nv + 1, dtype=f32) / f32(nv)).astype(f32)
        else:
            mcdf[1:] = (mcdf[1:] / self.marginal_func_int).astype(f32)
        self.marginal_cdf = mcdf


class EnvironmentMap:
    """src/textures/environment_map.jl:9-45; data (h, w, 3) in equal-area (octahedral) layout."""

    def __init__(self, data, rotation=None):
        self.data = np.ascontiguousarray(np.asarray(data, dtype=f32))
        self.rotation = np.eye(3, dtype=f32) if rotation is None else np.asarray(rotation, dtype=f32)  # row-major
        lum = (f32(0.212671) * self.data[..., 0] + f32(0.715160) * self.data[..., 1]
               + f32(0.072169) * self.data[..., 2]).astype(f32)
        self.distribution = Distribution2D(lum)


class EnvironmentLight(_Light):
    type = A.HK_LIGHT_ENVIRONMENT
    infinite = True

    def __init__(self, env_map, scale=(1, 1, 1)):
        self.env_map, self.scale = env_map, _rgb(scale)

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        L.env_map = scene._envmap_id(self)
        L.scale = 1.0
        return L


class DiffuseAreaLight(_Light):
    type = A.HK_LIGHT_DIFFUSE_AREA

    def __init__(self, vertices, normal, area, uv, Le, scale, two_sided):
        self.vertices, self.normal, self.area, self.uv = vertices, normal, area, uv
        self.Le, self.scale, self.two_sided = _rgb(Le), float(scale), bool(two_sided)

    def to_abi(self, scene):
        L = A.HkLight(type=self.type)
        L.scale = self.scale
        L.rgb[:] = self.Le
        L.v[:] = np.asarray(self.vertices, dtype=f32).reshape(-1).tolist()
        L.normal[:] = np.asarray(self.normal, dtype=f32).tolist()
        L.area = float(self.area)
        L.uv[:] = np.asarray(self.uv, dtype=f32).reshape(-1).tolist()
        L.two_sided = 1 if self.two_sided else 0
        return L


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
class Mesh:
    def __init__(self, positions, faces, normals=None, uvs=None):
        self.positions = np.asarray(positions, dtype=f32).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.uint32).reshape(-1, 3)
        self.normals = None if normals is None else np.asarray(normals, dtype=f32).reshape(-1, 3)
        self.uvs = None if uvs is None else np.asarray(uvs, dtype=f32).reshape(-1, 2)


def uv_sphere(center, radius, nu=64, nv=64):
    """Tessellated sphere (GeometryBasics.Tesselation(Sphere, n) stand-in): (nu) x (nv) vertex grid."""
    th = np.linspace(0, np.pi, nv, dtype=np.float64)
    ph = np.linspace(0, 2 * np.pi, nu, dtype=np.float64)
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(TH) * np.cos(PH), np.cos(TH), np.sin(TH) * np.sin(PH)], -1).reshape(-1, 3)
    p = np.asarray(center, dtype=np.float64) + radius * n
    uv = np.stack([PH / (2 * np.pi), TH / np.pi], -1).reshape(-1, 2)
    faces = []
    for i in range(nv - 1):
        for j in range(nu - 1):
            a, b, c, d = i * nu + j, i * nu + j + 1, (i + 1) * nu + j, (i + 1) * nu + j + 1
            faces.append((a, c, b))
            faces.append((b, c, d))
    return Mesh(p, faces, n, uv)


def rect3(origin, widths):
    """Axis-aligned box Rect3f(origin, widths) as 12 triangles with flat per-face normals."""
    o = np.asarray(origin, dtype=np.float64)
    w = np.asarray(widths, dtype=np.float64)
    P, N, UV, F = [], [], [], []
    axes = [(0, 1, 2), (1, 2, 0), (2, 0, 1)]
    for ax, u, v in axes:
        for side in (0, 1):
            base = o.copy()
            base[ax] += w[ax] * side
            du = np.zeros(3); du[u] = w[u]
            dv = np.zeros(3); dv[v] = w[v]
            n = np.zeros(3); n[ax] = 1.0 if side else -1.0
            i0 = len(P)
            for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                P.append(base + a * du + b * dv); N.append(n); UV.append((a, b))
            if side:
                F += [(i0, i0 + 1, i0 + 2), (i0, i0 + 2, i0 + 3)]
            else:
                F += [(i0, i0 + 2, i0 + 1), (i0, i0 + 3, i0 + 2)]
    return Mesh(P, F, N, UV)


def quad(p0, p1, p2, p3, normal=None):
    P = np.asarray([p0, p1, p2, p3], dtype=np.float64)
    n = np.cross(P[1] - P[0], P[3] - P[0])
    n = n / np.linalg.norm(n) if normal is None else np.asarray(normal, dtype=np.float64)
    return Mesh(P, [(0, 1, 2), (0, 2, 3)], [n] * 4, [(0, 0), (1, 0), (1, 1), (0, 1)])


# ------------------------------------------------------------------------------------------------
# Scene (src/scene.jl:21-151, src/scene-mesh.jl)
# ------------------------------------------------------------------------------------------------
class Scene:
    def __init__(self):
        self.meshes = []            # (mesh, transform, interface_idx, emission)
        self.instanced = False      # True: keep pushed meshes as instances of shared object-space meshes (two-level BVH) instead of flattening them
        self.materials = []         # unique Material objects
        self.interfaces = []        # (material_idx, inside_medium_idx, outside_medium_idx)
        self.media = []
        self.lights = []
        self._spectra = []
        self._envmaps = []
        self._synced = None

    # -- push! ------------------------------------------------------------------------------------
    def _index_of(self, lst, obj):
        for i, o in enumerate(lst):
            if o is obj:
                return i + 1
        lst.append(obj)
        return len(lst)

    def _spectrum_id(self, s):
        return self._index_of(self._spectra, s)

    def _texture_id(self, t):
        if not hasattr(self, "_textures"):
            self._textures = []
        return self._index_of(self._textures, t)

    def _envmap_id(self, light):
        return self._index_of(self._envmaps, light)

    def push_material(self, material):
        """push!(scene, material) -> index into media_interfaces (scene.jl; every material is wrapped in a
        MediumInterfaceIdx)."""
        if isinstance(material, MediumInterface):
            mat, inside, outside = material.material, material.inside, material.outside
        else:
            mat, inside, outside = material, None, None
        self._register_material(mat)
        mi = self._index_of(self.materials, mat)
        ii = 0 if inside is None else self._index_of(self.media, inside)
        oi = 0 if outside is None else self._index_of(self.media, outside)
        self.interfaces.append((mi, ii, oi))
        return len(self.interfaces)

    def _register_material(self, mat):
        """sub-materials of a MixMaterial are pushed before the mix itself (the reference needs their SetKeys first)"""
        if isinstance(mat, MixMaterial):
            for sub in mat.materials:
                self._register_material(sub)
        self._index_of(self.materials, mat)

    def set_key(self, mat):
        """SetKey (type_idx, vec_idx) of a pushed material, MultiTypeSet convention: type groups in first-push order"""
        types = []
        for m in self.materials:
            if type(m) not in types:
                types.append(type(m))
        same = [m for m in self.materials if type(m) is type(mat)]
        return types.index(type(mat)) + 1, [i for i, m in enumerate(same) if m is mat][0] + 1

    def push(self, obj, material=None, transform=None):
        """push!(scene, mesh, material; transform) / push!(scene, light)."""
        self._synced = None
        if isinstance(obj, _Light):
            self.lights.append(obj)
            return len(self.lights)
        assert isinstance(obj, Mesh) and material is not None
        idx = self.push_material(material)
        emission = material.emission if isinstance(material, MediumInterface) else None
        self.meshes.append((obj, None if transform is None else np.asarray(transform, dtype=np.float64), idx, emission))
        return len(self.meshes)

    # -- sync!: flatten to the C ABI arrays ----------------------------------------------------------
    def _sync_instanced(self):
        """scene.instanced = True: the meshes stay in object space (one copy per distinct Mesh object) and every push becomes an
        HkInstance (mesh, transform, medium interface) -- what the reference's Raycore TLAS holds (src/scene.jl:21-28, scene-mesh.jl:9-16)."""
        mesh_ids, pos, nrm, uvs, idx, meshes, insts = {}, [], [], [], [], [], []
        any_n = any(m.normals is not None for m, *_ in self.meshes)
        any_uv = any(m.uvs is not None for m, *_ in self.meshes)
        voff = toff = 0
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        n_world = 0
        for mesh, xf, iface, emission in self.meshes:
            if emission is not None:
                raise NotImplementedError("instanced meshes cannot be area lights (include/hikari_cuda.h: HkInstance); push emissive meshes into a scene with instanced = False")
            if id(mesh) not in mesh_ids:
                mesh_ids[id(mesh)] = len(meshes)
                pos.append(mesh.positions.astype(f32))
                if any_n:
                    nrm.append(np.full((len(mesh.positions), 3), np.nan, dtype=f32) if mesh.normals is None else mesh.normals.astype(f32))
                if any_uv:
                    uvs.append(np.zeros((len(mesh.positions), 2), dtype=f32) if mesh.uvs is None else mesh.uvs)
                idx.append(mesh.faces + np.uint32(voff))
                meshes.append((toff, len(mesh.faces)))
                voff += len(mesh.positions); toff += len(mesh.faces)
            M = np.eye(4) if xf is None else xf
            o2w = M[:3, :4].astype(np.float64)
            w2o = np.linalg.inv(np.vstack([o2w, [0, 0, 0, 1]]))[:3, :4]
            insts.append((mesh_ids[id(mesh)], iface, o2w.astype(f32), w2o.astype(f32)))
            P = mesh.positions.astype(np.float64)
            blo, bhi = P.min(0), P.max(0)
            corners = np.array([[(bhi if c & 1 else blo)[0], (bhi if c & 2 else blo)[1], (bhi if c & 4 else blo)[2]] for c in range(8)])
            wc = corners @ o2w[:, :3].T + o2w[:, 3]
            lo, hi = np.minimum(lo, wc.min(0)), np.maximum(hi, wc.max(0))
            n_world += len(mesh.faces)
        type_order = []
        for L in self.lights:
            if type(L) not in type_order:
                type_order.append(type(L))
        order = sorted(range(len(self.lights)), key=lambda i: (type_order.index(type(self.lights[i])), i))
        s = type("Synced", (), {})()
        s.positions = np.ascontiguousarray(np.concatenate(pos)) if pos else np.zeros((0, 3), f32)
        s.normals = np.ascontiguousarray(np.concatenate(nrm)) if any_n else None
        s.uvs = np.ascontiguousarray(np.concatenate(uvs).astype(f32)) if any_uv else None
        s.indices = np.ascontiguousarray(np.concatenate(idx).astype(np.uint32)) if idx else np.zeros((0, 3), np.uint32)
        s.tri_meta = None
        s.meshes, s.instances = meshes, insts
        s.n_world_tris, s.world_bounds = n_world, (lo, hi)
        s.lights = [self.lights[i] for i in order]
        self._synced = s
        return s

    def sync(self):
        if getattr(self, "instanced", False):
            return self._sync_instanced()
        pos, nrm, uvs, idx, meta = [], [], [], [], []
        any_n = any(m.normals is not None for m, *_ in self.meshes)
        any_uv = any(m.uvs is not None for m, *_ in self.meshes)
        area_lights = []
        voff = 0
        # flat light order: the reference's MultiTypeSet groups lights by type slot in first-push order
        type_order = []
        for L in self.lights:
            if type(L) not in type_order:
                type_order.append(type(L))
        for mesh, xf, iface, emission in self.meshes:
            P = mesh.positions.astype(np.float64)
            Nn = None if mesh.normals is None else mesh.normals.astype(np.float64)
            if xf is not None:
                P = P @ xf[:3, :3].T + xf[:3, 3]
                if Nn is not None:
                    Nn = Nn @ np.linalg.inv(xf[:3, :3])
                    Nn = Nn / np.linalg.norm(Nn, axis=1, keepdims=True)
            P32 = P.astype(f32)
            pos.append(P32)
            if any_n:
                nrm.append(np.full((len(P), 3), np.nan, dtype=f32) if Nn is None else Nn.astype(f32))
            if any_uv:
                uvs.append(np.zeros((len(P), 2), dtype=f32) if mesh.uvs is None else mesh.uvs)
            idx.append(mesh.faces + np.uint32(voff))
            m = np.zeros((len(mesh.faces), 3), dtype=np.uint32)
            m[:, 0] = iface
            m[:, 1] = np.arange(1, len(mesh.faces) + 1, dtype=np.uint32)
            if emission is not None:                       # register_face_area_lights!, scene-mesh.jl:100-138
                Le, scale, two_sided = emission
                Le = _rgb(Le)
                lum = f32(0.212671) * f32(Le[0]) + f32(0.715160) * f32(Le[1]) + f32(0.072169) * f32(Le[2])
                for fi, face in enumerate(mesh.faces):
                    vs = P32[face]
                    fuv = (mesh.uvs[face] if mesh.uvs is not None else np.array([[0, 0], [1, 0], [1, 1]], dtype=f32))
                    if lum < 1e-4:
                        continue
                    e1, e2 = (vs[1] - vs[0]).astype(f32), (vs[2] - vs[0]).astype(f32)
                    cp = np.cross(e1, e2).astype(f32)
                    ta = f32(np.sqrt(f32(cp[0] * cp[0] + cp[1] * cp[1]) + f32(cp[2] * cp[2])))
                    if ta < 1e-10:
                        continue
                    al = DiffuseAreaLight(vs, (cp / ta).astype(f32), f32(0.5) * ta, fuv, Le, scale, two_sided)
                    area_lights.append((al, len(meta), fi))
            meta.append(m)
            voff += len(P)
        lights = list(self.lights)
        if area_lights and DiffuseAreaLight not in type_order:
            type_order.append(DiffuseAreaLight)
        lights += [al for al, _, _ in area_lights]
        order = sorted(range(len(lights)), key=lambda i: (type_order.index(type(lights[i])), i))
        flat = [lights[i] for i in order]
        flat_of = {id(l): k + 1 for k, l in enumerate(flat)}
        for al, mi, fi in area_lights:
            meta[mi][fi, 2] = flat_of[id(al)]
        s = type("Synced", (), {})()
        s.positions = np.ascontiguousarray(np.concatenate(pos)) if pos else np.zeros((0, 3), f32)
        s.normals = np.ascontiguousarray(np.concatenate(nrm)) if any_n else None
        s.uvs = np.ascontiguousarray(np.concatenate(uvs).astype(f32)) if any_uv else None
        s.indices = np.ascontiguousarray(np.concatenate(idx).astype(np.uint32)) if idx else np.zeros((0, 3), np.uint32)
        s.tri_meta = np.ascontiguousarray(np.concatenate(meta).astype(np.uint32)) if meta else np.zeros((0, 3), np.uint32)
        s.lights = flat
        s.meshes, s.instances = [], []
        s.n_world_tris = len(s.indices)
        s.world_bounds = (s.positions.min(0), s.positions.max(0)) if len(s.positions) else (np.zeros(3), np.zeros(3))
        self._synced = s
        return s

    def triangle_count(self):
        """world-space triangles of the scene (every instance counted)"""
        s = self._synced or self.sync()
        return int(s.n_world_tris)

    def world_radius(self):
        s = self._synced or self.sync()
        lo, hi = s.world_bounds
        return float(np.linalg.norm(np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64)) / 2)


# ------------------------------------------------------------------------------------------------
# VolPath (src/integrators/volpath/volpath.jl)
# ------------------------------------------------------------------------------------------------
def compute_zsobol_params(samples_per_pixel, width, height):
    """src/sampler/sobol.jl:317-323"""
    log2_spp = int(math.ceil(math.log2(max(1, samples_per_pixel))))
    res_log2 = int(math.ceil(math.log2(max(width, height))))
    log4_spp = (log2_spp + 1) // 2
    return log2_spp, res_log2 + log4_spp


class Backend:
    """The product backend: libhikari_cuda.so behind the C ABI (include/hikari_cuda.h).  There is no CPU fallback: construction
    fails when the library or a CUDA device is missing.  (tests/oracle_backend.py derives the oracle's backend from this class to
    feed the checker the identical flattened scene; nothing in this module loads or calls the oracle.)"""
    prefix = "hk_"
    device_majorant = True        # majorant grids are built on the device from the uploaded voxels (HkMedium.majorant = NULL)

    def __init__(self, device=0):
        self.lib = A.load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.hk_create(device, C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError(f"hk_create failed with status {rc}: no usable CUDA device / library (there is no CPU fallback)")
        self._keep = []

    def _last_error(self):
        e = self.lib.hk_last_error(self.ctx)
        return e.decode() if e else ""

    def call(self, name, *args):
        rc = getattr(self.lib, self.prefix + name)(self.ctx, *args)
        if rc != 0:
            raise RuntimeError(f"{self.prefix}{name} failed ({rc}): {self._last_error()}")
        return rc

    def close(self):
        if self.ctx:
            getattr(self.lib, self.prefix + "destroy")(self.ctx)
            self.ctx = C.c_void_p()

    # ---- uploads --------------------------------------------------------------------------------
    def upload_tables(self):
        t = T.load_tables()
        scale, coeffs = T.get_srgb_table()
        keep = [np.ascontiguousarray(t[k]) for k in ("sobol_matrices", "cie_x", "cie_y", "cie_z", "d65_values")]
        keep += [scale, coeffs]
        ht = A.HkTables(keep[0].ctypes.data_as(A.c_u32p), _fp(keep[1]), _fp(keep[2]), _fp(keep[3]), _fp(keep[4]),
                        len(scale), _fp(scale), _fp(coeffs))
        self.call("upload_tables", C.byref(ht))

    def upload_scene(self, scene):
        s = scene._synced or scene.sync()
        marr = (A.HkMesh * max(1, len(s.meshes)))(*[A.HkMesh(a, b) for a, b in s.meshes])
        iarr = (A.HkInstance * max(1, len(s.instances)))()
        for k, (mi, iface, o2w, w2o) in enumerate(s.instances):
            iarr[k].mesh, iarr[k].medium_interface_idx = mi, iface
            iarr[k].object_to_world[:] = o2w.reshape(-1).tolist(); iarr[k].world_to_object[:] = w2o.reshape(-1).tolist()
        g = A.HkGeometry(_fp(s.positions), None if s.normals is None else _fp(s.normals), None,
                         None if s.uvs is None else _fp(s.uvs), s.indices.ctypes.data_as(A.c_u32p),
                         None if s.tri_meta is None else s.tri_meta.ctypes.data_as(A.c_u32p), len(s.positions), len(s.indices),
                         marr, len(s.meshes), iarr, len(s.instances))
        self.call("upload_geometry", C.byref(g))
        # materials first (registers spectra and texture ids), then spectra and textures, then the material upload
        scene._textures = []
        mats = (A.HkMaterial * max(1, len(scene.materials)))(*[m.to_abi(scene) for m in scene.materials])
        texs = (A.HkTexture * max(1, len(scene._textures)))(*[A.HkTexture(_fp(t.data), t.h, t.w, None if getattr(t, "alpha", None) is None else _fp(t.alpha))
                                                             for t in scene._textures])
        self.call("upload_textures", texs, len(scene._textures))
        lam = np.concatenate([sp.lambdas for sp in scene._spectra]) if scene._spectra else np.zeros(0, f32)
        val = np.concatenate([sp.values for sp in scene._spectra]) if scene._spectra else np.zeros(0, f32)
        offs = np.cumsum([0] + [len(sp.lambdas) for sp in scene._spectra]).astype(np.uint32)
        sp = A.HkSpectra(_fp(lam), _fp(val), offs.ctypes.data_as(A.c_u32p), len(scene._spectra))
        self.call("upload_spectra", C.byref(sp))
        ifs = (A.HkMediumInterface * max(1, len(scene.interfaces)))(*[A.HkMediumInterface(*t) for t in scene.interfaces])
        self.call("upload_materials", mats, len(scene.materials), ifs, len(scene.interfaces))
        # media
        keep = []
        if scene.media:
            med = (A.HkMedium * len(scene.media))(*[m.to_abi(keep, self.device_majorant) for m in scene.media])
            self.call("upload_media", med, len(scene.media))
        else:
            self.call("upload_media", None, 0)
        # lights (+ env maps + sampler)
        scene._envmaps = []
        labi = [L.to_abi(scene) for L in s.lights]
        if scene._envmaps:
            envs = []
            for L in scene._envmaps:
                em, d = L.env_map, L.env_map.distribution
                e = A.HkEnvMap()
                e.rgb, e.w, e.h = _fp(em.data), em.data.shape[1], em.data.shape[0]
                e.rotation[:] = em.rotation.T.reshape(-1).tolist()        # column-major like Julia's Mat3f
                e.scale_rgb[:] = L.scale
                e.conditional_func, e.conditional_cdf = _fp(d.conditional_func), _fp(d.conditional_cdf)
                e.conditional_func_int, e.marginal_func = _fp(d.conditional_func_int), _fp(d.marginal_func)
                e.marginal_cdf, e.marginal_func_int, e.nu, e.nv = _fp(d.marginal_cdf), float(d.marginal_func_int), d.nu, d.nv
                envs.append(e)
            earr = (A.HkEnvMap * len(envs))(*envs)
            self.call("upload_envmaps", earr, len(envs))
        n = len(labi)
        larr = (A.HkLight * max(1, n))(*labi)
        nodes = (A.HkLightBVHNode * max(1, 2 * n))()
        trails = np.zeros(max(1, n), dtype=np.uint32)
        inf = np.zeros(max(1, n), dtype=np.int32)
        nn, ni, nb = C.c_uint32(), C.c_uint32(), C.c_uint32()
        A.load_library().hk_host_build_light_sampler(larr, n, nodes, C.byref(nn), trails.ctypes.data_as(A.c_u32p),
                                                     inf.ctypes.data_as(A.c_i32p), C.byref(ni), C.byref(nb))
        smp = A.HkLightSampler(nodes, nn.value, trails.ctypes.data_as(A.c_u32p), inf.ctypes.data_as(A.c_i32p), ni.value, nb.value)
        self.call("upload_lights", larr, n, C.byref(smp))
        self.light_sampler_info = (nn.value, ni.value, nb.value)
        self._keep = [s, keep]

    def set_camera(self, camera):
        c = camera.to_abi()
        self.call("set_camera", C.byref(c))

    def set_filter(self, sampler_data):
        f = sampler_data.to_abi()
        self.call("set_filter", C.byref(f))

    def set_params(self, vp, width, height, sample_batch=0):
        sobol_spp = max(int(vp.samples_per_pixel), 4096)            # volpath.jl:475
        l2, nb4 = compute_zsobol_params(sobol_spp, width, height)
        p = A.HkRenderParams(width, height, vp.max_depth, vp.samples_per_pixel, 1 if vp.regularize else 0,
                             vp.max_component_value, 0, l2, nb4,
                             {"none": 0, "sorted": 1, "per_type": 2}[vp.material_coherence], sample_batch)
        self.call("set_params", C.byref(p))
        self.width, self.height = width, height

    def read_film(self, out_hw3):
        """framebuffer[py, px] (H, W, 3); the ABI writes the reference's (H, W) column-major RGB layout."""
        if isinstance(out_hw3, Film):                                # zero-copy: the film's storage is already (H,W) col-major
            assert out_hw3.resolution == (self.width, self.height)
            self.call("read_film", _fp(out_hw3._store))
            return
        buf = np.empty((self.width, self.height, 3), dtype=f32)     # column-major (H,W) == C-order (W,H)
        self.call("read_film", _fp(buf))
        out_hw3[...] = buf.transpose(1, 0, 2)

    def update_medium(self, medium_idx, medium):
        keep = []
        abi = medium.to_abi(keep, self.device_majorant)
        self.call("update_medium", medium_idx, C.byref(abi))

    def read_nanovdb(self, medium_idx):
        """the NanoVDB buffer of medium `medium_idx` as the device holds it (uploaded, or built on the device from the dense volume)"""
        n = C.c_uint64(0)
        self.call("read_nanovdb", medium_idx, None, 0, C.byref(n))
        out = np.empty(n.value, dtype=np.uint8)
        self.call("read_nanovdb", medium_idx, out.ctypes.data_as(A.c_u8p), out.size, C.byref(n))
        return out

    def read_majorant(self, medium_idx, res):
        """the majorant grid of medium `medium_idx` as the device holds it, [rz][ry][rx]"""
        out = np.empty((res[2], res[1], res[0]), dtype=f32)
        self.call("read_majorant", medium_idx, _fp(out), out.size)
        return out

    def read_film_async(self, film):
        """Enqueue finalize + device->host copy of the current film into a page-locked buffer of the film that no displayed or
        in-flight frame uses and return a handle at once (hk_read_film_async); wait_film(handle) makes that frame
        film.framebuffer.  At most four read-outs may be in flight (the library's four staging buffers)."""
        assert film.resolution == (self.width, self.height)
        store = film._acquire_store()
        ticket = C.c_int32(-1)
        self.call("read_film_async", _fp(store), C.byref(ticket))
        return (ticket.value, store)

    def wait_film(self, film, handle):
        ticket, store = handle
        self.call("read_film_wait", ticket)
        film._show(store)

    def read_accum(self):
        n = self.width * self.height
        rgb, w = np.empty((n, 3), dtype=f32), np.empty(n, dtype=f32)
        self.call("read_accum", _fp(rgb), _fp(w))
        return rgb, w


class VolPath:
    """volpath.jl:29-101.  Keyword-only like the reference: VolPath(samples=…, max_depth=…)."""

    def __init__(self, *, max_depth=8, samples=64, russian_roulette_depth=3, regularize=True,
                 material_coherence="none", max_component_value=10.0, filter=None, backend=None, sample_batch=0):
        assert material_coherence in ("none", "sorted", "per_type"), \
            "material_coherence must be :none, :sorted, :per_type"
        self.max_depth, self.samples_per_pixel = int(max_depth), int(samples)
        self.russian_roulette_depth, self.regularize = int(russian_roulette_depth), bool(regularize)
        self.material_coherence, self.max_component_value = material_coherence, float(max_component_value)
        self.filter = GaussianFilter() if filter is None else filter
        self.filter_sampler_data = FilterSamplerData(self.filter)
        self.backend = backend
        self.sample_batch = int(sample_batch)
        self.state = None          # (scene id, W, H, n_lights) the backend was prepared for

    def _prepare(self, scene, film, camera):
        w, h = film.resolution
        s = scene._synced or scene.sync()
        key = (id(s), w, h, len(s.lights))
        pkey = (w, h, self.max_depth, self.samples_per_pixel, self.regularize, self.max_component_value, self.material_coherence,
                self.sample_batch, id(self.filter_sampler_data))
        if self.backend is None:
            self.backend = Backend()
        if self.state != key:
            b = self.backend
            b.upload_tables()
            b.upload_scene(scene)
            self.state = key
            self._params_key = None
        if getattr(self, "_params_key", None) != pkey:      # the reference reads the VolPath fields on every render! call
            self.backend.set_filter(self.filter_sampler_data)
            self.backend.set_params(self, w, h, self.sample_batch)
            self._params_key = pkey
        self.backend.set_camera(camera)

    def update_material(self, scene, interface_idx, new_material):
        """update_material!(scene, idx, new_material), src/scene.jl:109-112: replace the material behind medium interface
        `interface_idx` (what push! returned) in place; the prepared backend gets the one struct, not the scene."""
        mi, _, _ = scene.interfaces[interface_idx - 1]
        assert not isinstance(new_material, MixMaterial), "replace the sub-materials of a mix, not the mix slot"
        old_material = scene.materials[mi - 1]
        scene.materials[mi - 1] = new_material
        # a MixMaterial that blends the replaced material must follow it: its ABI record holds the sub-material's index and
        # SetKey (type, index within type), which the mix hash consumes (mix-material.jl:114-158)
        mixes = [k for k, m in enumerate(scene.materials) if isinstance(m, MixMaterial) and any(sub is old_material for sub in m.materials)]
        for k in mixes:
            m = scene.materials[k]
            m.materials = tuple(new_material if sub is old_material else sub for sub in m.materials)
        if self.backend is not None and self.state is not None:
            abi = new_material.to_abi(scene)
            self.backend.call("update_material", mi, C.byref(abi))
            if type(new_material) is not type(old_material):
                mixes = [k for k, m in enumerate(scene.materials) if isinstance(m, MixMaterial)]      # set keys of every type may have shifted
            for k in mixes:
                abi = scene.materials[k].to_abi(scene)
                self.backend.call("update_material", k + 1, C.byref(abi))

    def update_medium(self, scene, medium_idx, new_medium):
        """The density-update path (build_majorant_grid! / build_rgb_majorant_grid!, media.jl:1498-1530, 1185-1240): replace medium
        `medium_idx` (1-based, what push!(scene.media) made it) in place; the prepared backend gets that one medium -- voxels up,
        majorant grid and empty-cell mask rebuilt on the device -- not the scene."""
        scene.media[medium_idx - 1] = new_medium
        if self.backend is not None and self.state is not None:
            self.backend.update_medium(medium_idx, new_medium)

    def wait_film(self, film, handle):
        """Block until the frame requested with render(..., read="async") is film.framebuffer."""
        self.backend.wait_film(film, handle)

    def clear(self):
        """clear!(vp), volpath.jl:108-113"""
        if self.state is not None:
            self.backend.call("clear")

    def render(self, scene, film, camera, count=1, read=True):
        """render!(vp, scene, film, camera) — `count` sample passes (volpath.jl:445-636)."""
        self._prepare(scene, film, camera)
        first = film.iteration_index + 1
        self.backend.call("render_samples", first, count)
        film.iteration_index += count
        if read == "async":        # pipelined progressive display: returns a handle for wait_film()
            return self.backend.read_film_async(film)
        if read:
            self.backend.read_film(film)

    def __call__(self, scene, film, camera):
        """(vp::VolPath)(scene, film, camera), volpath.jl:655-670"""
        film.iteration_index = 0
        self._prepare(scene, film, camera)
        self.clear()
        self.render(scene, film, camera, count=self.samples_per_pixel)
        return film.framebuffer

    def close(self):
        if self.backend is not None:
            self.backend.close()
