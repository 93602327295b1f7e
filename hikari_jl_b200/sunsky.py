"""sunsky_to_envlight (src/lights/sun_sky.jl:358-434): bake the Hosek-Wilkie spectral sky into a resolution^2 equal-area
EnvironmentMap and return it with a separate SunLight.  Host-side in the reference too (once per scene); the device only sees the
baked map.  Model: hosek_cook_config / hosek_cook_radiance (:19-125: quintic Bernstein interpolation over the cube root of the solar
elevation, bilinear in turbidity and ground albedo), hosek_radiance (:127-139), hosek_spectral_radiance (:165-186: linear between
the two neighbouring 40 nm bands).  Float64 like the reference; vectorised over the image.
Dataset: data/hosek_wilkie_spectral.npz (tools/extract_hosek.py; Hosek & Wilkie 2012-2013, BSD 3-clause)."""
import os

import numpy as np

from . import tables as T

f32 = np.float32
_DATA = None
D65_PHOTOMETRIC = 10567.0          # src/spectral/color.jl:16
CIE_Y_INTEGRAL = 106.856895        # color.jl:11


def _dataset():
    global _DATA
    if _DATA is None:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(T._NPZ)), "hosek_wilkie_spectral.npz"))
        _DATA = (z["configs"], z["radiances"])
    return _DATA


def _bernstein5(t, c):
    """_hosek_bernstein5, :6-14; c = the six control values (last axis)"""
    s = 1.0 - t
    return (s ** 5 * c[..., 0] + 5.0 * s ** 4 * t * c[..., 1] + 10.0 * s ** 3 * t ** 2 * c[..., 2] + 10.0 * s ** 2 * t ** 3 * c[..., 3]
            + 5.0 * s * t ** 4 * c[..., 4] + t ** 5 * c[..., 5])


def _cook(dataset, n, turbidity, albedo, solar_elevation):
    """hosek_cook_config (n = 9) / hosek_cook_radiance (n = 1): dataset laid out [albedo 0/1][turbidity 1..10][control point 0..5][n]"""
    it = int(np.clip(np.floor(turbidity), 1, 10))
    rem = turbidity - float(it)
    t = (solar_elevation / (np.pi / 2.0)) ** (1.0 / 3.0)
    d = dataset.reshape(2, 10, 6, n)
    ctrl = lambda a, k: np.moveaxis(d[a, k], 0, -1)            # [n][6]
    out = (1.0 - albedo) * (1.0 - rem) * _bernstein5(t, ctrl(0, it - 1)) + albedo * (1.0 - rem) * _bernstein5(t, ctrl(1, it - 1))
    if it < 10:
        out = out + (1.0 - albedo) * rem * _bernstein5(t, ctrl(0, it)) + albedo * rem * _bernstein5(t, ctrl(1, it))
    return out


class HosekState:
    """HosekState(turbidity, albedo, solar_elevation), :141-163"""

    def __init__(self, turbidity, albedo, solar_elevation):
        cfg, rad = _dataset()
        self.configs = np.stack([_cook(cfg[b], 9, turbidity, albedo, solar_elevation) for b in range(11)])       # [11][9]
        self.radiances = np.array([float(_cook(rad[b], 1, turbidity, albedo, solar_elevation)[0]) for b in range(11)])
        self.turbidity, self.albedo, self.elevation = turbidity, albedo, solar_elevation
        self.solar_radius = np.deg2rad(0.51) / 2.0


def hosek_radiance(config, theta, gamma):
    """:127-139; theta = angle from the zenith, gamma = angle from the sun (arrays)"""
    cg = np.cos(gamma)
    ct = np.maximum(np.cos(theta), 0.0)
    expM = np.exp(config[4] * gamma)
    rayM = cg * cg
    mieM = (1.0 + cg * cg) / ((1.0 + config[8] * config[8] - 2.0 * config[8] * cg) ** 1.5)
    zenith = np.sqrt(ct)
    return (1.0 + config[0] * np.exp(config[1] / (ct + 0.01))) * (config[2] + config[3] * expM + config[5] * rayM + config[6] * mieM + config[7] * zenith)


def hosek_spectral_radiance(state, theta, gamma, wavelength):
    """:165-186"""
    low = int(np.floor((wavelength - 320.0) / 40.0))
    if low < 0 or low >= 11:
        return np.zeros_like(np.asarray(theta, dtype=np.float64))
    interp = ((wavelength - 320.0) / 40.0) % 1.0
    val_low = hosek_radiance(state.configs[low], theta, gamma) * state.radiances[low]
    if interp < 1e-6:
        return val_low
    out = (1.0 - interp) * val_low
    if low + 1 < 11:
        out = out + interp * hosek_radiance(state.configs[low + 1], theta, gamma) * state.radiances[low + 1]
    return out


def _spectrum_to_xyz(lambdas, values):
    """:319-333: piecewise-linear spectrum (values [..., n]) against the 1 nm CIE tables, / CIE_Y_INTEGRAL"""
    t = T.load_tables()
    lam = 360.0 + np.arange(471, dtype=np.float64)
    idx = np.clip(np.searchsorted(lambdas, lam, side="right") - 1, 0, len(lambdas) - 2)
    w = np.clip((lam - lambdas[idx]) / (lambdas[idx + 1] - lambdas[idx]), 0.0, 1.0)          # clamps = the end-value extension (:294-301)
    s = (1.0 - w) * values[..., idx] + w * values[..., idx + 1]                              # [..., 471]
    return tuple((s * t[k].astype(np.float64)).sum(axis=-1) / CIE_Y_INTEGRAL for k in ("cie_x", "cie_y", "cie_z"))


def equal_area_square_to_sphere(u, v):
    """environment_map.jl:133-160 (float32 like the reference)"""
    u, v = np.asarray(u, dtype=f32), np.asarray(v, dtype=f32)
    uu, vv = f32(2) * u - f32(1), f32(2) * v - f32(1)
    up, vp = np.abs(uu), np.abs(vv)
    sd = f32(1) - (up + vp)
    d = np.abs(sd)
    r = f32(1) - d
    phi = np.where(r == 0, f32(1), (vp - up) / np.where(r == 0, f32(1), r) + f32(1)) * f32(np.pi / 4)
    z = np.copysign(f32(1) - r * r, sd)
    cx, sy = np.copysign(np.cos(phi), uu), np.copysign(np.sin(phi), vv)
    rc = r * np.sqrt(np.maximum(f32(2) - r * r, f32(0)))
    return np.stack([cx * rc, sy * rc, z], -1).astype(f32)


def sunsky_sky_data(direction, turbidity=2.5, ground_albedo=(0.3, 0.3, 0.3), ground_enabled=True, resolution=512):
    """the baked (resolution, resolution, 3) RGB image of sunsky_to_envlight, data[v, u] (:383-421)"""
    d = np.asarray(direction, dtype=f32)
    d = (d / np.sqrt((d * d).sum(dtype=f32))).astype(f32)
    elevation = float(np.arcsin(np.clip(d[2], f32(0), f32(1))))
    state = HosekState(float(f32(turbidity)), 0.5, elevation)
    n_lambda = 1 + (720 - 320) // 32
    lambdas = np.array([320.0 + i * (720.0 - 320.0) / (n_lambda - 1) for i in range(n_lambda)])
    c = ((np.arange(resolution, dtype=f32) + f32(1)) - f32(0.5)) / f32(resolution)            # (idx - 0.5) / resolution, idx 1-based
    U, V = np.meshgrid(c, c, indexing="xy")                                                    # data[v, u]
    wi = equal_area_square_to_sphere(U, V)
    theta = np.arccos(np.clip(wi[..., 2], f32(0), f32(1))).astype(np.float64)
    gamma = np.arccos(np.clip((wi * d).sum(axis=-1, dtype=f32), f32(-1), f32(1))).astype(np.float64)
    vals = np.stack([hosek_spectral_radiance(state, theta, gamma, lam) for lam in lambdas], axis=-1)
    x, y, z = (a.astype(f32) for a in _spectrum_to_xyz(lambdas, vals))
    rgb = np.stack([f32(3.2404542) * x - f32(1.5371385) * y - f32(0.4985314) * z,              # xyz_to_linear_srgb, color.jl:572-579
                    f32(-0.9692660) * x + f32(1.8760108) * y + f32(0.0415560) * z,
                    f32(0.0556434) * x - f32(0.2040259) * y + f32(1.0572252) * z], -1)
    sky = np.maximum(rgb, f32(0)).astype(f32)
    if ground_enabled:
        g = (np.asarray(ground_albedo, dtype=f32) * f32(0.3)).astype(f32)
        sky = np.where(wi[..., 2:3] <= 0, g, sky).astype(f32)
    return sky, d


def sunsky_to_envlight(direction, intensity=1.0, turbidity=2.5, ground_albedo=(0.3, 0.3, 0.3), ground_enabled=True, resolution=512):
    """-> (EnvironmentLight, SunLight), :358-434: env light scale = intensity / D65_PHOTOMETRIC, sun = 5 intensity (1, 0.95, 0.85)
    from -direction."""
    from . import host as H
    sky, d = sunsky_sky_data(direction, turbidity, ground_albedo, ground_enabled, resolution)
    s = float(f32(intensity) / f32(D65_PHOTOMETRIC))
    env = H.EnvironmentLight(H.EnvironmentMap(sky), scale=(s, s, s))
    k = float(f32(5) * f32(intensity))
    sun = H.SunLight((k, float(f32(k) * f32(0.95)), float(f32(k) * f32(0.85))), tuple(float(-v) for v in d))
    return env, sun
