"""The BASELINE.json configs restated as concrete synthetic scenes with the CURRENT
reference API shape (Scene / push / sync, VolPath(samples=…, max_depth=…)).

  cornell_smoke   test/volpath_integration.jl:8-95 (64x64, 4 spp, depth 4) — the reference's own smoke scene
  c1_triangle     examples/single_triangle_test.jl:12-90
  c1_spheres      examples/sphere_normals_test.jl:41-98
  c2_cat          examples/cat_scene.jl:39-130 (cat.obj asset absent -> procedural closed stand-in mesh)
  c3_many_lights  README.md:66-77 + 10 000 emissive triangles (analytic sky stand-in for the Hosek-Wilkie bake)
  c4_cloud        examples/bomex_cloud_example.jl:78-161 with a procedural cumulus field (NanoVDB or Grid)
  c5_instanced    instanced mixed-material field (triangle budget is a parameter; 50 M at full size)

Each function returns (scene, camera_factory) where camera_factory(film) -> PerspectiveCamera.
"""
import numpy as np

from . import host as H

f32 = np.float32


def _cam(eye, look, fov):
    return lambda film: H.PerspectiveCamera(eye, look, film, fov=fov, screen_window="aspect")


def cornell_smoke():
    white = H.MatteMaterial(Kd=(0.73, 0.73, 0.73))
    red = H.MatteMaterial(Kd=(0.65, 0.05, 0.05))
    green = H.MatteMaterial(Kd=(0.12, 0.45, 0.15))
    glass = H.GlassMaterial(Kr=1.0, Kt=1.0, index=1.5)
    fog = H.HomogeneousMedium(sigma_a=0.01, sigma_s=0.3, Le=0.0, g=0.3)
    glass_fog = H.MediumInterface(glass, inside=fog, outside=None)
    gold = H.ConductorMaterial(eta=(0.15557, 0.42415, 1.3831), k=(3.6024, 2.4721, 1.9155))
    s = H.Scene()
    half, box = 1.0, 2.0
    s.push(H.rect3((-half, 0, -half), (box, 0.01, box)), white)
    s.push(H.rect3((-half, 0, half - 0.01), (box, box, 0.01)), white)
    s.push(H.rect3((-half, 0, -half), (0.01, box, box)), red)
    s.push(H.rect3((half - 0.01, 0, -half), (0.01, box, box)), green)
    s.push(H.uv_sphere((-0.4, 0.4, 0.0), 0.35, 32, 32), glass_fog)
    s.push(H.uv_sphere((0.4, 0.35, 0.0), 0.3, 32, 32), gold)
    # PointLight(position, i::RGBSpectrum) -> scale = 1 (src/lights/point.jl:26-28)
    s.push(H.PointLight((15, 15, 15), (0.0, 1.8, 0.0), legacy_rgbspectrum=True, scale=1.0))
    s.sync()
    return s, _cam((0.0, 1.0, -3.5), (0.0, 1.0, 0.0), 40.0)


def textured_spheres(tess=24):
    """MatteMaterial with a textured Kd (TextureRef + bilinear eval_tex, texture-ref.jl:72-190): a checker floor, a striped sphere,
    a sphere whose texture has values outside [0, 1] (clamped after filtering), one reached through a MixMaterial, and a plain
    matte sphere on the constant-parameter path next to them."""
    rng = np.random.RandomState(3)
    yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    checker = np.where(((yy // 2 + xx // 2) % 2)[..., None] == 0, np.array([0.85, 0.8, 0.7]), np.array([0.15, 0.2, 0.45])).astype(np.float32)
    stripes = np.zeros((8, 32, 3), np.float32); stripes[..., 0] = (np.arange(32) % 4 < 2) * 0.9; stripes[..., 1] = np.linspace(0.1, 0.9, 8)[:, None]; stripes[..., 2] = 0.3
    wild = rng.uniform(-0.5, 1.6, size=(5, 7, 3)).astype(np.float32)
    s = H.Scene()
    floor = H.Mesh([(-5, -0.9, -5), (5, -0.9, -5), (5, -0.9, 5), (-5, -0.9, 5)], [(0, 2, 1), (0, 3, 2)],
                   normals=[(0, 1, 0)] * 4, uvs=[(-0.2, -0.2), (1.2, -0.2), (1.2, 1.2), (-0.2, 1.2)])       # uv beyond [0, 1]: indices clamp
    s.push(floor, H.MatteMaterial(Kd=H.Texture(checker)))
    s.push(H.uv_sphere((-1.6, 0.0, 0.0), 0.75, tess, tess), H.MatteMaterial(Kd=H.Texture(stripes), sigma=20.0))
    s.push(H.uv_sphere((0.0, 0.0, 0.0), 0.75, tess, tess), H.MatteMaterial(Kd=H.Texture(wild)))
    s.push(H.uv_sphere((1.6, 0.0, 0.0), 0.75, tess, tess),
           H.MixMaterial((H.MirrorMaterial(Kr=0.9), H.MatteMaterial(Kd=H.Texture(checker[:4, :6].copy()))), amount=1.0))
    s.push(H.uv_sphere((0.0, 1.3, 0.0), 0.4, tess, tess), H.MatteMaterial(Kd=(0.2, 0.7, 0.3)))
    d = np.array([-1.0, -1.5, -0.5])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.3, 0.35, 0.4)))
    s.sync()
    return s, _cam((0, 1.6, 4.5), (0, 0.2, 0), 40.0)


def alpha_foliage(fog=False, n_cards=24, seed=9):
    """Alpha-tested surfaces (intersection.jl:221-266, 349-372): a stack of "leaf" cards whose MatteMaterial.Kd texture carries an alpha
    plane (opaque discs on a transparent background, plus a fractional-alpha rim), over a matte floor, seen against the light so that
    camera rays, bounce rays AND shadow rays all meet several cards in a row.  One card sits behind a MixMaterial (alpha 1 there,
    whatever its texture says) and one is plain opaque.  fog=True puts the cards inside a homogeneous medium bounded by an index-1.5
    glass box, so shadow rays that pass a card ratio-track up to it (the medium is unchanged by an alpha pass-through)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    r = np.hypot(yy - 15.5, xx - 15.5)
    leaf = np.zeros((32, 32, 4), np.float32)
    leaf[..., 0] = 0.15 + 0.1 * (yy % 4 < 2); leaf[..., 1] = 0.55 + 0.2 * (xx % 8 < 4); leaf[..., 2] = 0.12
    leaf[..., 3] = np.clip((14.0 - r) / 4.0, 0.0, 1.0)                 # opaque core, a rim of fractional alpha, transparent corners
    half = leaf.copy(); half[..., :3] = (0.8, 0.3, 0.2); half[..., 3] = 0.5
    s = H.Scene()
    s.push(H.rect3((-5, -1.0, -5), (10, 0.1, 10)), H.MatteMaterial(Kd=(0.6, 0.6, 0.6)))
    leaf_mat, half_mat = H.MatteMaterial(Kd=H.Texture(leaf)), H.MatteMaterial(Kd=H.Texture(half), sigma=15.0)
    mixed = H.MixMaterial((H.MatteMaterial(Kd=H.Texture(half.copy())), H.MatteMaterial(Kd=(0.2, 0.2, 0.7))), amount=0.5)
    for i in range(n_cards):
        c = np.array([rng.uniform(-1.4, 1.4), rng.uniform(-0.5, 1.6), rng.uniform(-1.2, 1.2)])
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        R = H.rotation_matrix(rng.uniform(0, 360), ax).astype(np.float64)[:3, :3]
        hs = rng.uniform(0.35, 0.7)
        corners = [c + R @ np.array(v) * hs for v in ((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0))]
        n = R @ np.array((0.0, 0.0, 1.0))
        card = H.Mesh([tuple(v) for v in corners], [(0, 1, 2), (0, 2, 3)], normals=[tuple(n)] * 4, uvs=[(0, 0), (1, 0), (1, 1), (0, 1)])
        mat = mixed if i == 3 else (H.MatteMaterial(Kd=(0.7, 0.6, 0.2)) if i == 5 else (half_mat if i % 4 == 1 else leaf_mat))
        s.push(card, mat)
    if fog:
        medium = H.HomogeneousMedium(sigma_a=0.02, sigma_s=0.25, Le=0.0, g=0.2)
        s.push(H.rect3((-2.2, -0.85, -2.0), (4.4, 3.2, 4.0)), H.MediumInterface(H.GlassMaterial(Kr=0.0, Kt=1.0, index=1.0), inside=medium, outside=None))
    d = np.array([0.3, -1.0, 0.5])
    s.push(H.DirectionalLight((4, 4, 4), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.PointLight((20, 20, 20), (0.5, 2.6, -1.0), legacy_rgbspectrum=True, scale=1.0))
    s.push(H.AmbientLight((0.25, 0.3, 0.4)))
    s.sync()
    return s, _cam((0.0, 1.2, -5.0), (0.0, 0.4, 0.0), 40.0)


def vertex_color_meshes(tess=10):
    """VertexColorTexture as MatteMaterial.Kd (textures/basic.jl:43-46, texture-ref.jl:240-245): per-face corner colours interpolated
    with the hit's barycentrics, indexed by TriangleMeta.primitive_index -- on two meshes (so the face index restarts), one of them
    placed by a transform, next to a constant-colour sphere."""
    rng = np.random.RandomState(4)
    s = H.Scene()
    s.push(H.rect3((-5, -1.0, -5), (10, 0.1, 10)), H.MatteMaterial(Kd=(0.7, 0.7, 0.7)))
    for x in (-1.3, 1.3):
        m = H.uv_sphere((x, 0.2, 0.0), 0.9, tess, tess)
        fc = rng.uniform(-0.1, 1.2, size=(len(m.faces), 3, 3)).astype(np.float32)        # outside [0, 1]: clamped after interpolation
        if x < 0:
            s.push(m, H.MatteMaterial(Kd=H.VertexColorTexture(fc)))
        else:
            M = np.eye(4); M[:3, :3] = H.rotation_matrix(30.0, (0, 1, 0)).astype(np.float64)[:3, :3] * 0.9; M[:3, 3] = (0.1, 0.3, 0.2)
            s.push(m, H.MatteMaterial(Kd=H.VertexColorTexture(fc), sigma=10.0), transform=M)
    s.push(H.uv_sphere((0.0, 1.5, 0.3), 0.4, tess, tess), H.MatteMaterial(Kd=(0.2, 0.7, 0.3)))
    d = np.array([-0.6, -1.0, 0.8])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.3, 0.35, 0.4)))
    s.sync()
    return s, _cam((0, 1.6, -4.5), (0, 0.2, 0), 40.0)


def textured_parameters(tess=16, seed=6):
    """Textures on every kind of material parameter (eval_tex(textures, mat.<param>, tfc) at each read in spectral-eval.jl):
    RGB parameters (Mirror.Kr, Glass.Kr / Kt, Conductor eta / k, Coated* reflectance / albedo / transmittance, DiffuseTransmission
    reflectance / transmittance) and scalar ones (Matte sigma, Glass index, Conductor roughness, coating roughness / thickness / g),
    each on its own sphere or panel, plus constant-parameter neighbours on the cached-uplift path."""
    rng = np.random.RandomState(seed)
    def rgb_tex(h, w, lo, hi):
        return H.Texture(rng.uniform(lo, hi, size=(h, w, 3)).astype(np.float32))
    def f_tex(h, w, lo, hi):
        return H.Texture(rng.uniform(lo, hi, size=(h, w)).astype(np.float32))
    s = H.Scene()
    floor = H.Mesh([(-6, -0.9, -6), (6, -0.9, -6), (6, -0.9, 6), (-6, -0.9, 6)], [(0, 2, 1), (0, 3, 2)], normals=[(0, 1, 0)] * 4, uvs=[(0, 0), (1, 0), (1, 1), (0, 1)])
    s.push(floor, H.MatteMaterial(Kd=(0.6, 0.6, 0.6), sigma=f_tex(6, 6, 0.0, 40.0)))
    mats = [
        H.MirrorMaterial(Kr=rgb_tex(8, 8, 0.2, 1.0)),
        H.GlassMaterial(Kr=rgb_tex(4, 6, 0.5, 1.0), Kt=rgb_tex(6, 4, 0.4, 1.0), index=f_tex(5, 5, 1.3, 1.7)),
        H.ConductorMaterial(eta=rgb_tex(4, 4, 0.1, 1.5), k=rgb_tex(4, 4, 1.5, 4.0), roughness=f_tex(8, 8, 0.0, 0.4)),
        H.CoatedDiffuseMaterial(reflectance=rgb_tex(8, 8, 0.1, 0.9), roughness=(f_tex(4, 4, 0.0, 0.3), 0.1), thickness=f_tex(4, 4, 0.005, 0.05),
                                albedo=rgb_tex(4, 4, 0.0, 0.8), g=f_tex(4, 4, -0.5, 0.5), max_depth=8, n_samples=1),
        H.DiffuseTransmissionMaterial(reflectance=rgb_tex(6, 6, 0.0, 0.6), transmittance=rgb_tex(6, 6, 0.0, 0.6), scale=1.2),
        H.CoatedConductorMaterial(interface_roughness=f_tex(4, 4, 0.0, 0.2), reflectance=rgb_tex(8, 8, 0.3, 1.0), conductor_roughness=(0.05, f_tex(4, 4, 0.01, 0.2)),
                                  thickness=0.02, albedo=rgb_tex(4, 4, 0.0, 0.5), g=0.1),
        H.CoatedDiffuseTransmissionMaterial(reflectance=rgb_tex(4, 4, 0.1, 0.7), transmittance=rgb_tex(4, 4, 0.1, 0.7), roughness=0.1, albedo=0.0),
        H.ConductorMaterial(eta=(0.2, 0.9, 1.1), k=(3.9, 2.4, 2.2), roughness=0.05),          # constant parameters: the cached path next to the textured one
    ]
    for i, m in enumerate(mats):
        x, z = -3.0 + 2.0 * (i % 4), (-1.0 if i < 4 else 1.4)
        s.push(H.uv_sphere((x, 0.0, z), 0.8, tess, tess), m)
    d = np.array([-0.5, -1.0, 0.6])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.PointLight((30, 30, 30), (0.0, 3.5, -3.0), legacy_rgbspectrum=True, scale=1.0))
    s.push(H.AmbientLight((0.3, 0.35, 0.4)))
    s.sync()
    return s, _cam((0, 3.0, -7.0), (0, -0.2, 0), 40.0)


def textured_mix(tess=24, seed=9):
    """MixMaterial.amount as a Texture{Float32} (choose_material, mix-material.jl:183: amt = eval_tex(ctx, mix.amount, uv)): a
    floor whose mix runs from all-matte (amount <= 0) through the hashed blend to all-mirror (amount >= 1) along u, a sphere
    with a noisy amount, a nested mix whose inner amount is textured too, and a constant-amount neighbour."""
    rng = np.random.RandomState(seed)
    ramp = np.tile(np.linspace(-0.25, 1.25, 16, dtype=np.float32), (2, 1))          # [h, w]: varies along u, leaves [0, 1] at both ends
    noise = rng.uniform(-0.2, 1.2, size=(12, 12)).astype(np.float32)
    half = np.zeros((4, 4), np.float32); half[:2, :] = 1.0
    s = H.Scene()
    grey, mirror = H.MatteMaterial(Kd=(0.7, 0.7, 0.7)), H.MirrorMaterial(Kr=0.9)
    red, blue, gold = H.MatteMaterial(Kd=(0.8, 0.2, 0.2)), H.MatteMaterial(Kd=(0.2, 0.3, 0.8)), H.Gold(roughness=0.1)
    floor = H.Mesh([(-5, -0.9, -5), (5, -0.9, -5), (5, -0.9, 5), (-5, -0.9, 5)], [(0, 2, 1), (0, 3, 2)], normals=[(0, 1, 0)] * 4, uvs=[(0, 0), (1, 0), (1, 1), (0, 1)])
    s.push(floor, H.MixMaterial((grey, mirror), amount=H.Texture(ramp)))
    inner = H.MixMaterial((red, gold), amount=H.Texture(half))
    s.push(H.uv_sphere((-1.8, 0.0, 0.0), 0.85, tess, tess), H.MixMaterial((blue, gold), amount=H.Texture(noise)))
    s.push(H.uv_sphere((0.0, 0.0, 0.0), 0.85, tess, tess), H.MixMaterial((inner, blue), amount=H.Texture(noise.T.copy())))
    s.push(H.uv_sphere((1.8, 0.0, 0.0), 0.85, tess, tess), H.MixMaterial((red, mirror), amount=0.5))
    d = np.array([-0.6, -1.0, 0.5])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.25, 0.3, 0.35)))
    s.sync()
    return s, _cam((0, 2.2, -5.5), (0, -0.2, 0), 40.0)


def rgb_nebula(res=(20, 16, 12)):
    """An RGBGridMedium (media.jl:1002-1456) inside an index-1 boundary: two coloured, partly emissive blobs over a matte floor."""
    nx, ny, nz = res
    x, y, z = np.meshgrid((np.arange(nx) + 0.5) / nx, (np.arange(ny) + 0.5) / ny, (np.arange(nz) + 0.5) / nz, indexing="ij")
    blob = lambda c, r: np.clip(1.0 - np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / r, 0.0, 1.0)
    b1, b2 = blob((0.35, 0.5, 0.5), 0.33), blob((0.68, 0.55, 0.45), 0.28)
    sig_s = (b1[..., None] * np.array([6.0, 3.0, 1.0]) + b2[..., None] * np.array([1.0, 3.0, 7.0])).astype(np.float32)
    sig_a = (b1[..., None] * np.array([0.2, 0.6, 1.5]) + b2[..., None] * np.array([1.2, 0.4, 0.1])).astype(np.float32)
    Le = ((b2 > 0.6)[..., None] * np.array([0.5, 1.0, 3.0])).astype(np.float32)
    lo = (-0.8, 0.1, -0.5)
    med = H.RGBGridMedium(sigma_a_grid=sig_a, sigma_s_grid=sig_s, Le_grid=Le, sigma_scale=1.5, Le_scale=1.0, g=0.3,
                          bounds=(lo, (0.8, 1.3, 0.5)), majorant_res=(5, 4, 3))
    s = H.Scene()
    s.push(H.rect3((-3, -0.1, -3), (6, 0.1, 6)), H.MatteMaterial(Kd=(0.6, 0.6, 0.6)))
    s.push(H.rect3(lo, (1.6, 1.2, 1.0)), H.MediumInterface(H.GlassMaterial(Kr=0.0, Kt=1.0, index=1.0), inside=med))
    d = np.array([-0.4, -1.0, 0.6])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.15, 0.18, 0.25)))
    s.sync()
    return s, _cam((0.0, 0.9, -3.2), (0.0, 0.6, 0.0), 40.0)


def c1_triangle():
    s = H.Scene()
    tri = H.Mesh([(-1, -0.5, 0), (1, -0.5, 0), (0, 1, 0)], [(0, 1, 2)],
                 normals=[(0, 0, 1), (0.7, 0, 0.714), (0, 0.7, 0.714)], uvs=[(0, 0), (1, 0), (0.5, 1)])
    s.push(tri, H.MatteMaterial(Kd=(0.8, 0.8, 0.8)))
    s.push(H.DirectionalLight((2, 2, 2), (0, 0, -1), legacy_rgbspectrum=True))
    s.sync()
    return s, _cam((0, 0, 3), (0, 0, 0), 50.0)


def c1_spheres(tess=64):
    s = H.Scene()
    s.push(H.rect3((-5, -1, -5), (10, 0.1, 10)), H.MatteMaterial(Kd=(0.7, 0.7, 0.7)))
    for x, kd in ((-1.5, (0.8, 0.2, 0.2)), (0.0, (0.2, 0.8, 0.2)), (1.5, (0.2, 0.2, 0.8))):
        s.push(H.uv_sphere((x, 0.5, 0.0), 0.8, tess, tess), H.MatteMaterial(Kd=kd))
    d = np.array([-1.0, -1.5, -0.5])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.sync()
    return s, _cam((0, 1.5, 4), (0, 0.5, 0), 40.0)


def mix_spheres(tess=32):
    """c1_spheres with MixMaterials (mix-material.jl): a matte/mirror blend, a nested mix, and the amount = 0 / 1 ends."""
    s = H.Scene()
    red, green, blue = H.MatteMaterial(Kd=(0.8, 0.2, 0.2)), H.MatteMaterial(Kd=(0.2, 0.8, 0.2)), H.MatteMaterial(Kd=(0.2, 0.2, 0.8))
    mirror, gold = H.MirrorMaterial(Kr=0.9), H.Gold(roughness=0.1)
    inner = H.MixMaterial((green, gold), amount=0.35)
    s.push(H.rect3((-5, -1, -5), (10, 0.1, 10)), H.MixMaterial((H.MatteMaterial(Kd=(0.7, 0.7, 0.7)), mirror), amount=0.25))
    s.push(H.uv_sphere((-1.5, 0.5, 0.0), 0.8, tess, tess), H.MixMaterial((red, mirror), amount=0.5))
    s.push(H.uv_sphere((0.0, 0.5, 0.0), 0.8, tess, tess), H.MixMaterial((inner, blue), amount=0.6))
    s.push(H.uv_sphere((1.5, 0.5, 0.0), 0.8, tess, tess), H.MixMaterial((blue, red), amount=1.0))
    s.push(H.uv_sphere((0.0, 1.9, 0.0), 0.4, tess, tess), H.MixMaterial((green, red), amount=0.0))
    d = np.array([-1.0, -1.5, -0.5])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.2, 0.25, 0.3)))
    s.sync()
    return s, _cam((0, 1.5, 4), (0, 0.5, 0), 40.0)


def coated_conductor_spheres(tess=32):
    """The four closed-form cases of CoatedConductorMaterial (spectral-eval.jl:2877-3418) side by side: coating smooth / rough
    x conductor smooth / rough, in eta-k (RGB and measured-spectrum) and reflectance mode, one with an absorbing layer, and
    one reached through a MixMaterial (amount = 1: always its second material)."""
    s = H.Scene()
    au = H.Gold()
    s.push(H.rect3((-5, -1, -5), (10, 0.1, 10)), H.MatteMaterial(Kd=(0.7, 0.7, 0.7)))
    s.push(H.uv_sphere((-2.1, 0.5, 0.0), 0.65, tess, tess),
           H.CoatedConductorMaterial(interface_roughness=0.0, conductor_roughness=0.0, reflectance=(0.95, 0.64, 0.54)))
    s.push(H.uv_sphere((-0.7, 0.5, 0.0), 0.65, tess, tess),
           H.CoatedConductorMaterial(interface_roughness=0.0, conductor_roughness=0.15, conductor_eta=(0.143, 0.374, 1.442),
                                     conductor_k=(3.983, 2.385, 1.603)))
    s.push(H.uv_sphere((0.7, 0.5, 0.0), 0.65, tess, tess),
           H.CoatedConductorMaterial(interface_roughness=0.3, conductor_roughness=0.0, conductor_eta=au.eta, conductor_k=au.k,
                                     albedo=(0.6, 0.8, 0.6), thickness=0.2))
    s.push(H.uv_sphere((2.1, 0.5, 0.0), 0.65, tess, tess),
           H.MixMaterial((H.MatteMaterial(Kd=(0.8, 0.2, 0.2)),
                          H.CoatedConductorMaterial(interface_roughness=(0.2, 0.05), conductor_roughness=(0.1, 0.3),
                                                    reflectance=(0.9, 0.1, 0.1), interface_eta=1.33)), amount=1.0))
    d = np.array([-1.0, -1.5, -0.5])
    s.push(H.DirectionalLight((3, 3, 3), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.PointLight((40, 40, 40), (1.0, 3.0, 3.0)))
    s.push(H.AmbientLight((0.4, 0.5, 0.6)))
    s.sync()
    return s, _cam((0, 1.5, 5), (0, 0.4, 0), 40.0)


def coated_difftrans_panels(tess=24):
    """CoatedDiffuseTransmissionMaterial (spectral-eval.jl:2340-2840): three back-lit thin panels (smooth coating, rough
    coating, rough coating + absorbing layer) and a sphere, over a matte floor."""
    s = H.Scene()
    s.push(H.rect3((-5, -1, -5), (10, 0.1, 10)), H.MatteMaterial(Kd=(0.7, 0.7, 0.7)))
    kws = (dict(roughness=0.0), dict(roughness=0.3), dict(roughness=0.1, albedo=(0.8, 0.4, 0.2), g=0.3, thickness=0.1))
    for k, kw in enumerate(kws):
        x = -1.6 + 1.6 * k
        s.push(H.rect3((x - 0.6, -0.9, 0.0), (1.2, 1.8, 0.02)),
               H.CoatedDiffuseTransmissionMaterial(reflectance=(0.25, 0.5, 0.2), transmittance=(0.3, 0.6, 0.15), **kw))
    s.push(H.uv_sphere((0.0, 1.6, 0.3), 0.5, tess, tess),
           H.CoatedDiffuseTransmissionMaterial(reflectance=(0.6, 0.3, 0.3), transmittance=(0.3, 0.2, 0.5), roughness=0.05))
    s.push(H.PointLight((60, 60, 60), (0.5, 2.0, -3.0)))        # behind the panels
    d = np.array([-0.3, -1.0, -1.0])
    s.push(H.DirectionalLight((2, 2, 2), d / np.linalg.norm(d), legacy_rgbspectrum=True))
    s.push(H.AmbientLight((0.3, 0.35, 0.4)))
    s.sync()
    return s, _cam((0, 1.0, 5), (0, 0.3, 0), 40.0)


def blob_mesh(center, radius, n=256, seed=3):
    """Closed procedural stand-in for cat.obj: a sphere displaced by a few low-frequency harmonics."""
    m = H.uv_sphere((0, 0, 0), 1.0, n, n)
    p = m.positions.astype(np.float64)
    rng = np.random.RandomState(seed)
    disp = np.zeros(len(p))
    for _ in range(6):
        k = rng.randint(1, 5, size=3)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        disp += 0.08 * np.sin(k[0] * p[:, 0] * 3 + ph[0]) * np.sin(k[1] * p[:, 1] * 3 + ph[1]) * np.sin(k[2] * p[:, 2] * 3 + ph[2])
    p = p * (1.0 + disp)[:, None] * np.array([0.7, 1.0, 1.2])
    pos = p * radius + np.asarray(center, dtype=np.float64)
    # smooth normals from face normals
    f = m.faces
    fn = np.cross(pos[f[:, 1]] - pos[f[:, 0]], pos[f[:, 2]] - pos[f[:, 0]])
    vn = np.zeros_like(pos)
    for k in range(3):
        np.add.at(vn, f[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    vn = np.where(ln > 0, vn / np.maximum(ln, 1e-30), m.normals)
    return H.Mesh(pos, f, vn, m.uvs)


def c2_cat(cat_tess=256, sphere_tess=64):
    s = H.Scene()
    s.push(H.PointLight((1, 1, 1), (3, 3, -1)))
    s.push(H.PointLight((5, 5, 5), (-3, 2, 0)))
    s.push(H.AmbientLight((0.5, 0.7, 1.0)))
    s.push(blob_mesh((0.0, -0.9, 1.5), 0.6, cat_tess), H.MatteMaterial(Kd=(0.8, 0.6, 0.4)))
    s.push(H.rect3((-5, -1.5, -2), (10, 0.01, 10)), H.MatteMaterial(Kd=(0.3, 0.5, 0.3)))
    s.push(H.rect3((-5, -1.5, 8), (10, 5, 0.01)), H.ConductorMaterial(reflectance=(0.8, 0.6, 0.5), roughness=0.05))
    s.push(H.rect3((-5, -1.5, -2), (0.01, 5, 10)), H.MatteMaterial(Kd=(0.7, 0.7, 0.8)))
    s.push(H.uv_sphere((-2, -1.5 + 0.8, 2), 0.8, sphere_tess, sphere_tess), H.ConductorMaterial(reflectance=(0.9, 0.9, 0.9), roughness=0.02))
    s.push(H.uv_sphere((2, -1.5 + 0.6, 1), 0.6, sphere_tess, sphere_tess), H.ConductorMaterial(reflectance=(0.3, 0.6, 0.9), roughness=0.3))
    glass = H.GlassMaterial(Kr=(0.98, 1.0, 0.98), Kt=(0.98, 1.0, 0.98), index=1.5)
    s.push(H.uv_sphere((-0.8, -1.5 + 0.5, 0.5), 0.5, sphere_tess, sphere_tess), glass)
    s.push(H.uv_sphere((0.8, -1.5 + 0.4, 0.3), 0.4, sphere_tess, sphere_tess), glass)
    s.sync()
    return s, _cam((0, -0.9, -2.5), (0, -0.9, 10), 45.0)


def analytic_sky(res=512, sun_dir=(1, 2, 9), turbidity=3.0):
    """Stand-in for sunsky_to_envlight (src/lights/sun_sky.jl:358-434, Hosek-Wilkie bake on the Julia host):
    a smooth analytic sky in the same 512^2 equal-area layout.  Only the light's data differs from the
    reference config; the device code path (Distribution2D sampling, equal-area mapping, illuminant uplift) is
    identical."""
    sd = np.asarray(sun_dir, dtype=np.float64)
    sd = sd / np.linalg.norm(sd)
    u = (np.arange(res) + 0.5) / res
    U, V = np.meshgrid(u, u, indexing="xy")          # data[v, u]
    # equal-area square -> sphere (same mapping as environment_map.jl:133-160), float64 host code
    uu, vv = 2 * U - 1, 2 * V - 1
    up, vp = np.abs(uu), np.abs(vv)
    sdist = 1 - (up + vp)
    r = 1 - np.abs(sdist)
    phi = np.where(r == 0, 1.0, (vp - up) / np.where(r == 0, 1, r) + 1.0) * np.pi / 4
    z = np.copysign(1 - r * r, sdist)
    cx, sy = np.copysign(np.cos(phi), uu), np.copysign(np.sin(phi), vv)
    rc = r * np.sqrt(2 - r * r)
    d = np.stack([cx * rc, sy * rc, z], -1)
    cos_g = np.clip(d @ sd, -1, 1)
    up_c = np.clip(d[..., 2], 0, 1)
    zenith = np.array([0.25, 0.45, 0.95])
    horizon = np.array([0.85, 0.9, 1.0])
    sky = horizon + (zenith - horizon) * (up_c[..., None] ** 0.5)
    glow = (0.5 * (1 + cos_g)) ** (64 / turbidity)
    sky = sky * (0.6 + 1.4 * glow[..., None]) + np.array([1.0, 0.85, 0.6]) * (glow[..., None] ** 8) * 4
    sky = np.where(d[..., 2:3] < 0, sky * 0.05, sky)     # ground_enabled=false: dark lower hemisphere
    return sky.astype(f32), sd


def c3_many_lights(n_emitters=10000, sphere_tess=128, seed=1234, sky_model="analytic"):
    """sky_model = "hosek": the reference's own sunsky_to_envlight bake (hikari_jl_b200/sunsky.py); "analytic" (default, what the
    round-1 C3 numbers were measured with): the smooth stand-in above in the same layout."""
    s = H.Scene()
    if sky_model == "hosek":
        from . import sunsky
        sky, sd = sunsky.sunsky_sky_data((1, 2, 9), turbidity=3.0, ground_enabled=False, resolution=512)
        sd = sd.astype(np.float64)
    else:
        sky, sd = analytic_sky(512)
    s.push(H.EnvironmentLight(H.EnvironmentMap(sky), scale=tuple([float(f32(1.0) / f32(10567.0))] * 3)))
    s.push(H.SunLight((5.0, 4.75, 4.25), -sd))
    s.push(H.uv_sphere((0, 0, 0), 1.0, sphere_tess, sphere_tess), H.GlassMaterial(index=1.5))
    s.push(H.rect3((-2, -2, -1), (4, 4, 0.01)), H.Gold(roughness=0.01))
    rng = np.random.RandomState(seed)
    c = np.stack([rng.uniform(-2, 2, n_emitters), rng.uniform(-2, 2, n_emitters), rng.uniform(0, 2, n_emitters)], -1)
    nrm = rng.normal(size=(n_emitters, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    t = np.cross(nrm, rng.normal(size=(n_emitters, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(nrm, t)
    e = 0.02
    Le = rng.uniform(5, 50, size=(n_emitters, 3))
    dark = H.MatteMaterial(Kd=(0.0, 0.0, 0.0))
    # one mesh per emitter colour bucket would be slow to flatten in Python: push in groups of equal Le (quantised)
    groups = 16
    q = np.floor((Le - 5) / 45 * groups).clip(0, groups - 1).astype(int)
    key = q[:, 0] * groups * groups + q[:, 1] * groups + q[:, 2]
    for k in np.unique(key):
        idx = np.nonzero(key == k)[0]
        P = np.concatenate([c[idx] - 0.5 * e * t[idx] - 0.289 * e * b[idx], c[idx] + 0.5 * e * t[idx] - 0.289 * e * b[idx],
                            c[idx] + 0.577 * e * b[idx]], 0)
        n = len(idx)
        F = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], -1)
        le = tuple(float(v) for v in (5 + (np.array([k // (groups * groups), (k // groups) % groups, k % groups]) + 0.5) * 45 / groups))
        s.push(H.Mesh(P, F), H.MediumInterface(dark, emission=(le, 1.0, True)))
    s.sync()
    return s, _cam((0, -6.0, 2.5), (0, 0, 0.2), 40.0)


def cumulus_field(shape=(256, 256, 128), occupancy=0.03, seed=7, max_extinction=620.0):
    """Procedural BOMEX-like cloud field: thresholded fBm, scaled so max extinction ~ 620 (bomex_cloud_example.jl:50)."""
    rng = np.random.RandomState(seed)
    nx, ny, nz = shape
    field = np.zeros(shape, dtype=np.float64)
    amp = 1.0
    for octave in range(5):
        cells = 2 ** (octave + 2)
        coarse = rng.normal(size=(cells + 1, cells + 1, max(2, cells // 2) + 1))
        xs = np.linspace(0, cells, nx, endpoint=False); ys = np.linspace(0, cells, ny, endpoint=False)
        zs = np.linspace(0, max(2, cells // 2), nz, endpoint=False)
        xi, yi, zi = xs.astype(int), ys.astype(int), zs.astype(int)
        fx, fy, fz = (xs - xi)[:, None, None], (ys - yi)[None, :, None], (zs - zi)[None, None, :]
        def g(a, b, c):
            return coarse[np.ix_(xi + a, yi + b, zi + c)]
        v = ((g(0, 0, 0) * (1 - fx) + g(1, 0, 0) * fx) * (1 - fy) + (g(0, 1, 0) * (1 - fx) + g(1, 1, 0) * fx) * fy) * (1 - fz) + \
            ((g(0, 0, 1) * (1 - fx) + g(1, 0, 1) * fx) * (1 - fy) + (g(0, 1, 1) * (1 - fx) + g(1, 1, 1) * fx) * fy) * fz
        field += amp * v
        amp *= 0.5
    zprof = np.exp(-((np.linspace(0, 1, nz) - 0.45) / 0.25) ** 2)[None, None, :]
    field = field * zprof
    thr = np.quantile(field, 1 - occupancy)
    dens = np.clip(field - thr, 0, None)
    dens = dens / dens.max() * max_extinction
    return dens.astype(f32)


def c4_cloud(shape=(256, 256, 128), medium_kind="nanovdb", majorant_res=(64, 64, 64)):
    s = H.Scene()
    dens = cumulus_field(shape)
    lo, hi = (-0.6, 0.3, -0.6), (0.6, 1.5, 0.6)
    if medium_kind == "nanovdb":
        med = H.NanoVDBMedium(dens, bounds=(lo, hi), sigma_a=0.0, sigma_s=1.0, g=0.877, majorant_res=majorant_res)
    else:
        med = H.GridMedium(dens, sigma_a=0.0, sigma_s=1.0, g=0.877, bounds=(lo, hi), majorant_res=majorant_res)
    boundary = H.MediumInterface(H.GlassMaterial(Kr=0.0, Kt=1.0, index=1.0), inside=med, outside=None)
    s.push(H.rect3(lo, tuple(h - l for l, h in zip(lo, hi))), boundary)
    s.push(H.rect3((-5, -0.01, -5), (10, 0.01, 10)), H.MatteMaterial(Kd=(0.4, 0.4, 0.4)))
    s.push(H.rect3((-3, 0, -3), (0.01, 4, 6)), H.MatteMaterial(Kd=(0.5, 0.5, 0.55)))
    s.push(H.rect3((3, 0, -3), (0.01, 4, 6)), H.MatteMaterial(Kd=(0.5, 0.5, 0.55)))
    s.push(H.AmbientLight((0.03, 0.07, 0.23)))
    s.push(H.DirectionalLight((2.6, 2.5, 2.3), (-0.5826, -0.7660, -0.2717)))
    s.sync()
    return s, _cam((0, 1, -3.5), (0, 0.9, 0), 40.0)


def c5_instanced(n_instances=1000, base_tess=160, seed=11, instanced=False):
    """n_instances copies (jittered grid, random rotations) of a base mesh with (base_tess-1)^2*2 triangles;
    materials round-robin over the six in-scope types.  Full size: 1000 x ~50 000 = 50 M triangles.
    instanced=True keeps them as instances of ONE object-space mesh (HkGeometry.instances, two-level BVH); False flattens them."""
    s = H.Scene()
    s.instanced = instanced
    sky, sd = analytic_sky(512)
    s.push(H.EnvironmentLight(H.EnvironmentMap(sky), scale=tuple([float(f32(1.0) / f32(10567.0))] * 3)))
    s.push(H.SunLight((5.0, 4.75, 4.25), -sd))
    mats = [H.MatteMaterial(Kd=(0.7, 0.5, 0.3)), H.GlassMaterial(index=1.5), H.Gold(roughness=0.05),
            H.CoatedDiffuseMaterial(reflectance=(0.4, 0.45, 0.35), roughness=0.1),
            H.ThinDielectricMaterial(eta=1.5), H.DiffuseTransmissionMaterial(reflectance=(0.4, 0.3, 0.2), transmittance=(0.3, 0.4, 0.3))]
    base = blob_mesh((0, 0, 0), 0.35, base_tess, seed=5)
    rng = np.random.RandomState(seed)
    side = int(np.ceil(np.sqrt(n_instances)))
    s.push(H.rect3((-side * 0.5 - 1, -0.45, -side * 0.5 - 1), (side + 2, 0.05, side + 2)), H.MatteMaterial(Kd=(0.5, 0.5, 0.5)))
    for i in range(n_instances):
        gx, gz = i % side, i // side
        ang = rng.uniform(0, 2 * np.pi)
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        R = np.eye(4)
        R[:3, :3] = H.rotation_matrix(np.degrees(ang), ax).astype(np.float64)
        Tm = np.eye(4)
        Tm[:3, 3] = (gx - side / 2 + rng.uniform(-0.1, 0.1), rng.uniform(0.0, 0.3), gz - side / 2 + rng.uniform(-0.1, 0.1))
        s.push(base, mats[i % len(mats)], transform=Tm @ R)
    s.sync()
    return s, _cam((0, side * 0.35, -side * 0.75), (0, 0, 0), 40.0)
