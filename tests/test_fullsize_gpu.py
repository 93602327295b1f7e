"""Full-size GPU checks (BASELINE.json configs at their real resolutions) through size-independent properties:
bitwise determinism and independence of the number of samples in flight, the multi-GPU partition identity, and
closest-hit parity of the rays the integrator ACTUALLY traces (all bounce depths) against the brute-force oracle.
The oracle cannot render these sizes in seconds; it only answers the sampled ray batches."""
import ctypes as C

import numpy as np
import pytest

from hikari_jl_b200 import _abi as A, host as H, scenes
from util import fp, f32
import oracle_backend

pytestmark = pytest.mark.gpu


def _render(scene, camf, res, depth, first, count, batch):
    film = H.Film(res)
    vp = H.VolPath(samples=4096, max_depth=depth, sample_batch=batch)
    vp._prepare(scene, film, camf(film)); vp.clear()
    vp.backend.call("render_samples", first, count)
    vp.backend.read_film(film)
    return vp, film.framebuffer.copy()


def _last_rays_vs_oracle(vp, scene, n_pixels, n_check, brute, seed=0):
    """Closest-hit records left by the last pass vs the oracle on the very same rays."""
    n_slots = n_pixels
    rays = np.zeros((n_slots, 8), f32); hits = np.zeros((n_slots, 4), f32)
    assert vp.backend.lib.hk_test_read_rays(vp.backend.ctx, fp(rays), fp(hits), n_slots) == 0
    sel = np.random.RandomState(seed).choice(n_slots, size=n_check, replace=False)
    r = np.ascontiguousarray(rays[sel]); h = hits[sel]
    ok = oracle_backend.make_backend()
    ok.upload_tables(); ok.upload_scene(scene)
    ho = np.zeros((n_check, 4), f32)
    oracle_backend.lib().ok_trace_closest(ok.ctx, fp(r), n_check, fp(ho), 1 if brute else 0)
    ok.close()
    prim_c, prim_o = h.view(np.uint32)[:, 1], ho.view(np.uint32)[:, 1]
    assert np.array_equal(prim_c, prim_o), f"{(prim_c != prim_o).sum()} of {n_check} primitive ids differ"
    hit = prim_o != 0
    assert np.array_equal(h.view(np.uint32)[hit], ho.view(np.uint32)[hit]), "t / barycentrics differ bitwise"
    finite = np.isfinite(r[:, :6]).all(axis=1)
    return int(hit.sum()), int((~finite).sum())


def test_c2_full_resolution_properties():
    scene, camf = scenes.c2_cat(256, 64)
    res, depth, n = (1920, 1080), 12, 1920 * 1080
    vp4, img4 = _render(scene, camf, res, depth, 1, 4, 4)          # 4 samples in one pass
    n_hit, n_bad = _last_rays_vs_oracle(vp4, scene, n, 20000, brute=True)
    assert n_hit > 1000
    vp4.close()
    vp1, img1 = _render(scene, camf, res, depth, 1, 4, 1)          # 4 passes of 1 sample
    vp1.close()
    assert np.isfinite(img4).all() and img4.max() > 0
    assert np.array_equal(img4.view(np.uint32), img1.view(np.uint32)), "image depends on the number of samples in flight"
    vpa, imga = _render(scene, camf, res, depth, 1, 4, 0)          # automatic batch, second run: determinism
    vpa.close()
    assert np.array_equal(img4.view(np.uint32), imga.view(np.uint32)), "render is not deterministic"
    # multi-GPU partition identity: ranks 0/1 of 2 render samples {1,3} and {2,4}; accumulators add up to the full film
    acc = []
    for r in range(2):
        f2 = H.Film(res); v2 = H.VolPath(samples=4096, max_depth=depth)
        v2._prepare(scene, f2, camf(f2)); v2.clear()
        v2.backend.call("render_samples_strided", r + 1, 2, 2)
        acc.append(v2.backend.read_accum()); v2.close()
    rgb = acc[0][0] + acc[1][0]; w = acc[0][1] + acc[1][1]
    img = (rgb / np.maximum(w, 1e-30)[:, None]).reshape(res[1], res[0], 3)
    np.testing.assert_allclose(img, img4, rtol=2e-4, atol=2e-6)


def test_c3_4k_many_lights_rays_and_batching():
    scene, camf = scenes.c3_many_lights(10000, 128)
    res, depth, n = (3840, 2160), 12, 3840 * 2160
    vp2, img2 = _render(scene, camf, res, depth, 1, 2, 2)
    n_hit, n_bad = _last_rays_vs_oracle(vp2, scene, n, 20000, brute=True)
    assert n_hit > 100
    vp2.close()
    vp1, img1 = _render(scene, camf, res, depth, 1, 2, 1)
    vp1.close()
    assert np.isfinite(img2).all() and img2.max() > 0
    assert np.array_equal(img2.view(np.uint32), img1.view(np.uint32))


def test_c5_two_million_triangles_rays():
    """Mixed materials (glass, conductor, coated diffuse, thin dielectric, diffuse transmission): degenerate BSDF samples
    produce invalid continuation rays here; they must miss on both paths and must not stall the traversal."""
    scene, camf = scenes.c5_instanced(40, 160)
    res, depth, n = (1920, 1080), 8, 1920 * 1080
    vp, img = _render(scene, camf, res, depth, 1, 2, 1)
    n_hit, n_bad = _last_rays_vs_oracle(vp, scene, n, 40000, brute=False, seed=1)      # oracle BVH2
    n_hit2, _ = _last_rays_vs_oracle(vp, scene, n, 400, brute=True, seed=2)            # brute force over 2 M triangles
    assert n_hit > 1000 and n_hit2 > 10
    stats = A.HkStats(); vp.backend.lib.hk_stats(vp.backend.ctx, C.byref(stats))
    assert stats.last_render_ms < 2000.0, "a pass over 2 M triangles took seconds: invalid rays are walking the whole tree again"
    vp.close()
    assert np.isfinite(img[np.isfinite(img)]).all() and np.nanmax(img) > 0


def test_c5_thousand_instances_two_level_bvh():
    """C5 at its real size: 1 000 instances of a ~50 000-triangle mesh = 50 M world triangles, kept as instances (HkGeometry.instances,
    top-level BVH over the instances + one bottom-level BVH): closest hits of the rays the integrator actually traces -- primitive
    id, t, barycentrics -- bit-exact against the oracle's instanced traversal (per-mesh BVH2 and brute force), and the BVH is tens of
    MB instead of the 3 GB of the flattened scene."""
    scene, camf = scenes.c5_instanced(1000, 160, instanced=True)
    assert scene.triangle_count() == 1000 * 2 * 159 * 159 + 12
    res, depth, n = (3840, 2160), 8, 3840 * 2160
    vp, img = _render(scene, camf, res, depth, 1, 1, 1)
    stats = A.HkStats(); vp.backend.lib.hk_stats(vp.backend.ctx, C.byref(stats))
    assert stats.bvh_bytes < 64 << 20, f"two-level BVH should be tens of MB, is {stats.bvh_bytes / 2**20:.0f} MB"
    n_hit, n_bad = _last_rays_vs_oracle(vp, scene, n, 40000, brute=False, seed=1)
    n_hit2, _ = _last_rays_vs_oracle(vp, scene, n, 300, brute=True, seed=2)      # brute force: 300 rays x 50 M triangle tests
    assert n_hit > 1000 and n_hit2 > 10
    assert stats.last_render_ms < 2000.0
    vp.close()
    assert np.isfinite(img[np.isfinite(img)]).all() and np.nanmax(img) > 0


def test_c4_4k_cloud_rays_and_batching():
    """C4 at its real size (256 x 256 x 128 NanoVDB cloud, 4K, depth 32): the closest hits the trace stage leaves are bit-exact
    against the oracle, and the image is bitwise independent of the number of samples in flight."""
    scene, camf = scenes.c4_cloud((256, 256, 128), "nanovdb", (64, 64, 64))
    res, depth, n = (3840, 2160), 32, 3840 * 2160
    vp2, img2 = _render(scene, camf, res, depth, 1, 2, 2)
    n_hit, n_bad = _last_rays_vs_oracle(vp2, scene, n, 20000, brute=True)
    assert n_hit > 100
    vp2.close()
    vp1, img1 = _render(scene, camf, res, depth, 1, 2, 1)
    vp1.close()
    assert np.isfinite(img2).all() and img2.max() > 0
    assert np.array_equal(img2.view(np.uint32), img1.view(np.uint32))


def _oracle_rows(scene, camf, res, depth, step, offset):
    """sample 1 of image rows py with py % step == offset, rendered by the oracle at the full resolution (ok_set_row_subset)"""
    film = H.Film(res)
    vp = H.VolPath(samples=4096, max_depth=depth, backend=oracle_backend.make_backend())
    vp._prepare(scene, film, camf(film)); vp.clear()
    oracle_backend.lib().ok_set_row_subset(vp.backend.ctx, step, offset)
    vp.backend.call("render_samples", 1, 1)
    vp.backend.read_film(film)
    img = film.framebuffer.copy()
    rays = oracle_backend.lib().ok_rays_traced(vp.backend.ctx)
    vp.close()
    return img[offset::step], rays


@pytest.mark.parametrize("name", ["C3", "C4", "C5"])
def test_full_size_pixels_against_the_oracle(name):
    """Pixels of the FULL-SIZE configurations against the oracle: the CUDA path renders one sample of the whole 4K frame (1 000
    instances / 10 002 lights / the 256x256x128 NanoVDB cloud at depth 32), the oracle renders the same sample for every 270th image
    row at the same resolution (same scene, camera, sampler state: a pixel's sample stream does not depend on its neighbours), and those
    rows must agree bit for bit."""
    if name == "C3":
        (scene, camf), depth = scenes.c3_many_lights(10000, 128), 12
    elif name == "C4":
        (scene, camf), depth = scenes.c4_cloud((256, 256, 128), "nanovdb", (64, 64, 64)), 32
    else:
        (scene, camf), depth = scenes.c5_instanced(1000, 160, instanced=True), 8
    res, step, offset = (3840, 2160), 270, 133
    vp, img = _render(scene, camf, res, depth, 1, 1, 1)
    vp.close()
    rows_o, _ = _oracle_rows(scene, camf, res, depth, step, offset)
    rows_c = img[offset::step]
    assert rows_c.shape == rows_o.shape == (8, 3840, 3)
    assert np.isfinite(rows_o).all() and rows_o.max() > 0
    same = (rows_c.view(np.uint32) == rows_o.view(np.uint32)).all(axis=2)
    print(f"{name} full size: {same.mean():.6f} of {same.size} pixels bit-identical; mean {rows_c.mean():.6f} vs {rows_o.mean():.6f}")
    assert same.all(), f"{(~same).sum()} of {same.size} full-size pixels differ from the oracle"
