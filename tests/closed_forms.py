"""Closed-form radiometry checks shared by the oracle tests (tests/test_oracle_golden.py, CPU) and the CUDA tests
(tests/test_parity_gpu.py, -m gpu): each function renders small scenes through the back end `make_backend()` returns (None = the
product's CUDA back end) and asserts the analytic answer."""
import numpy as np

from hikari_jl_b200 import host as H
from hikari_jl_b200 import scenes

f32 = np.float32


def check_physical_known_answers(make_backend):
    """Closed-form radiometry the whole path has to reproduce, whatever the reading of the Julia source (the CUDA path is bit-identical
    to the oracle, so these pin it as well):
      * Lambert's law: a matte surface lit head-on by a directional light of irradiance E shows E rho / pi; under a point light of
        intensity I it shows rho / pi * I cos(theta) / r^2;
      * a matte sphere in a uniform environment shows rho x the environment (light sampling + BSDF sampling + their MIS weights sum to
        one) -- and under an AmbientLight it shows rho (3/2 - ln 5 / 8): the reference gives escaped rays a light pdf of 0 for every
        light type but EnvironmentLight (lights.jl:450-458) while its light samples are MIS-weighted (a quirk kept bit for bit);
      * furnace: a clear glass sphere (Kr = Kt = 1) in a uniform environment is invisible, a mirror of reflectance Kr shows Kr;
      * Beer-Lambert: an absorbing slab of thickness d in front of a uniform background shows exp(-sigma_a d)."""
    def render(s, cam, res=(32, 32), spp=64, depth=8):
        film = H.Film(res)
        vp = H.VolPath(samples=spp, max_depth=depth, backend=make_backend())
        img = vp(s, film, cam(film)).copy()
        vp.close()
        return img
    near = scenes._cam((0, 0, -4), (0, 0, 0), 30.0)
    far = scenes._cam((0, 0, -40), (0, 0, 0), 0.2)                 # a pencil of parallel rays at the sphere's pole
    env = lambda: H.EnvironmentLight(H.EnvironmentMap(np.ones((8, 8, 3), f32)), scale=(1 / 10567.0,) * 3)      # ~1 after the D65 normalisation
    # Lambert's law
    for rho in (0.5, 0.18):
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 64, 64), H.MatteMaterial(Kd=(rho,) * 3))
        s.push(H.DirectionalLight((1, 1, 1), (0, 0, 1), legacy_rgbspectrum=True)); s.sync()
        c = render(s, far, depth=2)[14:18, 14:18].mean()
        assert abs(c / (rho / np.pi) - 1) < 3e-3, (rho, c, rho / np.pi)
    # inverse-square law and the cosine: a point light of intensity I at height h above a matte floor, seen below it and off to the side
    for I, h, x in ((3.0, 2.0, 0.0), (3.0, 2.0, 1.5), (1.0, 0.7, 0.0)):
        s = H.Scene(); s.push(H.Mesh([(-5, 0, -5), (5, 0, -5), (5, 0, 5), (-5, 0, 5)], [(0, 2, 1), (0, 3, 2)]), H.MatteMaterial(Kd=(0.5,) * 3))
        s.push(H.PointLight((I, I, I), (0, h, 0))); s.sync()
        c = render(s, scenes._cam((x + 20.0, 20.0, 0.0), (x, 0, 0), 0.05), res=(16, 16), depth=2).mean()
        r2 = h * h + x * x
        assert abs(c / (0.5 / np.pi * I * (h / np.sqrt(r2)) / r2) - 1) < 4e-3, (I, h, x, c)
    # uniform environment: MIS weights sum to one
    for rho in (0.25, 0.8):
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 48, 48), H.MatteMaterial(Kd=(rho,) * 3)); s.push(env()); s.sync()
        img = render(s, near, depth=2, spp=256)
        bg = img[0:3, 0:3].mean()
        assert abs(bg - 1) < 5e-3, bg
        assert abs(img[12:20, 12:20].mean() / bg / rho - 1) < 1.5e-2, (rho, img[12:20, 12:20].mean() / bg)
    # AmbientLight: the reference's escaped-ray pdf of 0
    quirk = 1.5 - np.log(5.0) / 8.0
    for rho in (0.5, 0.8):
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 48, 48), H.MatteMaterial(Kd=(rho,) * 3)); s.push(H.AmbientLight((1, 1, 1))); s.sync()
        img = render(s, near, depth=2, spp=256)
        assert abs(img[12:20, 12:20].mean() / img[0:3, 0:3].mean() / (rho * quirk) - 1) < 1.5e-2, (rho, img[12:20, 12:20].mean())
    # furnace
    s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 48, 48), H.GlassMaterial(Kr=1.0, Kt=1.0, index=1.5)); s.push(H.AmbientLight((1, 1, 1))); s.sync()
    img = render(s, near, depth=40)
    assert abs(img[10:22, 10:22].mean() / img[0:3, 0:3].mean() - 1) < 1e-2
    s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 48, 48), H.MirrorMaterial(Kr=(0.5,) * 3)); s.push(H.AmbientLight((1, 1, 1))); s.sync()
    img = render(s, near, depth=10)
    assert abs(img[12:20, 12:20].mean() / img[0:3, 0:3].mean() - 0.5) < 5e-3
    # Beer-Lambert
    wide = scenes._cam((0, 0, -40), (0, 0, 0), 3.0)
    for sa, dz in ((1.0, 0.5), (2.0, 0.75), (0.25, 2.0)):
        s = H.Scene()
        med = H.HomogeneousMedium(sigma_a=(sa,) * 3, sigma_s=(0, 0, 0), g=0.0)
        s.push(H.rect3((-2, -2, -dz / 2), (4, 4, dz)), H.MediumInterface(H.GlassMaterial(Kr=0.0, Kt=1.0, index=1.0), inside=med))
        s.push(H.AmbientLight((1, 1, 1))); s.sync()
        t = render(s, wide, depth=8, spp=512).mean() / 1.0006      # (the ambient background renders as 1.0006)
        assert abs(t / np.exp(-sa * dz) - 1) < 1e-2, (sa, dz, t, np.exp(-sa * dz))
    # ... and through heterogeneous-medium code paths holding a constant density rho_d: exp(-sigma_a rho_d d), whatever the majorant
    # grid resolution (delta tracking with null collisions where the majorant exceeds the density: the padded NanoVDB leaves)
    for kind, dens, sa, dz, mres in (("grid", 2.0, 0.5, 0.5, (4, 4, 4)), ("nanovdb", 1.5, 1.0, 0.75, (8, 8, 8)), ("nanovdb", 0.5, 2.0, 1.0, (1, 1, 1))):
        d = np.full((12, 12, 10), dens, f32)
        lo, hi = (-2.0, -2.0, -dz / 2), (2.0, 2.0, dz / 2)
        if kind == "grid":
            med = H.GridMedium(d, sigma_a=sa, sigma_s=0.0, g=0.0, bounds=(lo, hi), majorant_res=mres)
        else:
            med = H.NanoVDBMedium(d, bounds=(lo, hi), sigma_a=sa, sigma_s=0.0, g=0.0, majorant_res=mres)
        s = H.Scene()
        s.push(H.rect3(lo, (4, 4, dz)), H.MediumInterface(H.GlassMaterial(Kr=0.0, Kt=1.0, index=1.0), inside=med))
        s.push(H.AmbientLight((1, 1, 1))); s.sync()
        t = render(s, wide, depth=8, spp=512)[8:24, 8:24].mean() / 1.0006
        # NanoVDB values sit at voxel centres and are trilinearly blended with the background (0) across the outermost half voxel at
        # either end of the ray: the two ramps take a quarter of a voxel off the optical depth (nanovdb.jl:400-469)
        tau = sa * dens * dz * ((1 - 0.25 / d.shape[2]) if kind == "nanovdb" else 1.0)
        assert abs(t / np.exp(-tau) - 1) < 1.5e-2, (kind, dens, sa, dz, t, np.exp(-tau))


def check_material_closed_forms(make_backend):
    """More closed forms, per material: a smooth conductor at normal incidence reflects ((eta-1)^2 + k^2) / ((eta+1)^2 + k^2) (the
    complex Fresnel term, grey eta / k so the uplift is exact); a closed DiffuseTransmission sphere (reflectance R, transmittance T) in
    a furnace shows R + T^2 / (1 - R) (= 1 when R + T = 1: the inner radiance solves I = R I + T); a ThinDielectric sphere seen along
    its axis shows R^2 + T^4 / (1 - R^2) with R = R0 + T0^2 R0 / (1 - R0^2) -- NOT 1: the reference multiplies a specular sample's
    f = R / |cos| into the throughput without dividing by the probability of having chosen it (surface-eval.jl:438-441,
    spectral-eval.jl:2019-2034), kept bit for bit."""
    def render(s, cam, res=(32, 32), spp=64, depth=8):
        film = H.Film(res)
        vp = H.VolPath(samples=spp, max_depth=depth, backend=make_backend())
        img = vp(s, film, cam(film)).copy()
        vp.close()
        return img
    near = scenes._cam((0, 0, -4), (0, 0, 0), 30.0)
    far = scenes._cam((0, 0, -40), (0, 0, 0), 0.2)
    ambient = 1.0006                                              # what AmbientLight((1, 1, 1)) renders as
    for eta, k in ((0.5, 2.0), (1.5, 0.0), (2.0, 3.0)):
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 64, 64), H.ConductorMaterial(eta=(eta,) * 3, k=(k,) * 3, roughness=0.0))
        s.push(H.AmbientLight((1, 1, 1))); s.sync()
        got = render(s, far, depth=4)[14:18, 14:18].mean() / ambient
        want = ((eta - 1) ** 2 + k * k) / ((eta + 1) ** 2 + k * k)
        assert abs(got / want - 1) < 2e-3, (eta, k, got, want)
    for R, T in ((0.5, 0.5), (0.3, 0.3)):
        want = R + T * T / (1 - R)
        # (an EnvironmentLight: under an AmbientLight the light samples are over-counted, test_physical_known_answers_on_the_oracle)
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 48, 48), H.DiffuseTransmissionMaterial(reflectance=R, transmittance=T))
        s.push(H.EnvironmentLight(H.EnvironmentMap(np.ones((8, 8, 3), f32)), scale=(1 / 10567.0,) * 3)); s.sync()
        img = render(s, near, depth=40, spp=128)
        got = img[12:20, 12:20].mean() / img[0:3, 0:3].mean()
        assert abs(got / want - 1) < 1.5e-2, (R, T, got, want)
    for eta in (1.5, 2.0):
        s = H.Scene(); s.push(H.uv_sphere((0, 0, 0), 1.0, 128, 128), H.ThinDielectricMaterial(eta=eta)); s.push(H.AmbientLight((1, 1, 1))); s.sync()
        got = render(s, far, depth=30, spp=256)[12:20, 12:20].mean() / ambient
        R0 = ((eta - 1) / (eta + 1)) ** 2; T0 = 1 - R0
        R = R0 + T0 * T0 * R0 / (1 - R0 * R0); T = 1 - R
        want = R * R + T ** 4 / (1 - R * R)
        assert abs(got / want - 1) < 5e-3, (eta, got, want)


def check_area_light_closed_form(make_backend):
    """A small square emitter of radiance L over a matte floor: the floor point below it shows rho / pi * L * integral(cos^2 / r^2 dA)
    -- light-BVH selection of the emissive triangles, their area -> solid-angle pdf, the light sample and the emissive hit of the
    BSDF-sampled ray with their MIS weights, all in one number.  L is taken from the same render (the emitter seen directly): the
    reference clamps an area light's RGB to [0, 1] before the uplift (arealight_Le -> uplift_rgb, diffuse-area.jl:54-66,
    rgb2spec.jl:83-87), so Le = 60 and Le = 100 both emit the unit spectrum."""
    def render(s, cam, res=(16, 16), spp=512, depth=2):
        film = H.Film(res)      # (no firefly clamp: the emitter seen directly is ~130 RGB units, max_component_value defaults to 10)
        vp = H.VolPath(samples=spp, max_depth=depth, max_component_value=1.0e6, backend=make_backend())
        img = vp(s, film, cam(film)).copy()
        vp.close()
        return img
    rho = 0.5
    seen = []
    for side, h, Le in ((0.2, 2.0, 100.0), (0.4, 3.0, 60.0), (1.0, 1.5, 2.0)):
        s = H.Scene()
        s.push(H.Mesh([(-5, 0, -5), (5, 0, -5), (5, 0, 5), (-5, 0, 5)], [(0, 2, 1), (0, 3, 2)]), H.MatteMaterial(Kd=(rho,) * 3))
        a = side / 2
        s.push(H.Mesh([(-a, h, -a), (a, h, -a), (a, h, a), (-a, h, a)], [(0, 1, 2), (0, 2, 3)]),
               H.MediumInterface(H.MatteMaterial(Kd=(0, 0, 0)), emission=((Le, Le, Le), 1.0, True)))
        s.sync()
        floor = render(s, scenes._cam((20.0, 20.0, 0.0), (0, 0, 0), 0.05)).mean(axis=(0, 1))
        direct = render(s, scenes._cam((0.0, h + 20.0, 0.001), (0, h, 0), 0.02), spp=16).mean(axis=(0, 1))      # the emitter from above
        xs = (np.arange(400) + 0.5) / 400 * side - a
        X, Z = np.meshgrid(xs, xs)
        r2 = X * X + Z * Z + h * h
        form = np.sum((h * h / r2) / r2) * (side / 400) ** 2
        got = floor / direct
        assert np.allclose(got, rho / np.pi * form, rtol=1.5e-2), (side, h, Le, got, rho / np.pi * form)
        seen.append(direct)
    assert np.allclose(seen[0], seen[1], rtol=1e-3) and np.allclose(seen[0], seen[2], rtol=1e-3), "Le >= 1 is clamped to the unit spectrum"


def check_environment_map_closed_form(make_backend):
    """A matte surface (normal n) under an environment that is 1 on the hemisphere around an axis a and 0 elsewhere receives
    E = pi (1 + n.a) / 2, so it shows rho (1 + n.a) / 2: the equal-area octahedral mapping in both directions, the luminance
    Distribution2D (sampling and pdf, with the 4 pi Jacobian), the bilinear look-up of escaped rays and the MIS between the two have to
    agree for that to come out -- for axes towards, away from, across and oblique to the normal."""
    def dirs(res):                                              # environment_map.jl:133-160, texel centres -> directions
        u = (np.arange(res) + 0.5) / res
        U, V = np.meshgrid(u, u, indexing="xy")
        uu, vv = 2 * U - 1, 2 * V - 1
        up, vp = np.abs(uu), np.abs(vv)
        sd = 1 - (up + vp)
        r = 1 - np.abs(sd)
        phi = np.where(r == 0, 1.0, (vp - up) / np.where(r == 0, 1, r) + 1.0) * np.pi / 4
        z = np.copysign(1 - r * r, sd)
        cx, sy = np.copysign(np.cos(phi), uu), np.copysign(np.sin(phi), vv)
        rc = r * np.sqrt(2 - r * r)
        return np.stack([cx * rc, sy * rc, z], -1)
    rho, d = 0.5, dirs(256)
    n = np.array([0.0, 0.0, -1.0])                              # the pole of the sphere the camera looks at
    far = scenes._cam((0, 0, -40), (0, 0, 0), 0.2)
    for a in ((0, 0, -1), (0, 0, 1), (0, 1, 0), (0.6, 0, -0.8), (0, 0.6, -0.8), (0.6, 0, 0.8), (-0.6, 0, -0.8)):
        a = np.array(a, float)
        env = np.repeat((d @ a > 0).astype(f32)[..., None], 3, axis=2)
        s = H.Scene()
        s.push(H.uv_sphere((0, 0, 0), 1.0, 64, 64), H.MatteMaterial(Kd=(rho,) * 3))
        s.push(H.EnvironmentLight(H.EnvironmentMap(env), scale=(1 / 10567.0,) * 3)); s.sync()
        film = H.Film((16, 16))
        vp = H.VolPath(samples=256, max_depth=2, backend=make_backend())
        got = vp(s, film, far(film)).mean()
        vp.close()
        assert abs(got - rho * (1 + n @ a) / 2) < 4e-3, (a, got, rho * (1 + n @ a) / 2)
