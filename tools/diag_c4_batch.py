"""Diagnostic (GPU): where does the C4 image depend on the number of samples in flight?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hikari_jl_b200 import host as H, scenes


def render(scene, camf, res, depth, first, count, batch):
    film = H.Film(res)
    vp = H.VolPath(samples=4096, max_depth=depth, sample_batch=batch)
    vp._prepare(scene, film, camf(film)); vp.clear()
    vp.backend.call("render_samples", first, count)
    vp.backend.read_film(film)
    img = film.framebuffer.copy()
    vp.close()
    return img


def diff(a, b, tag):
    d = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
    ys, xs = np.nonzero(d)
    print(f"{tag}: {d.sum()} differing pixels of {d.size}; first: {list(zip(ys[:8].tolist(), xs[:8].tolist()))}", flush=True)
    for y, x in list(zip(ys[:4], xs[:4])):
        print("   ", y, x, a[y, x], b[y, x], flush=True)


scene, camf = scenes.c4_cloud((256, 256, 128), "nanovdb", (64, 64, 64))
for res in ((480, 270), (1920, 1080), (3840, 2160)):
    for depth in (32,):
        a2 = render(scene, camf, res, depth, 1, 2, 2)
        a1 = render(scene, camf, res, depth, 1, 2, 1)
        a1b = render(scene, camf, res, depth, 1, 2, 1)
        a2b = render(scene, camf, res, depth, 1, 2, 2)
        diff(a2, a1, f"{res} depth {depth} batch2 vs batch1")
        diff(a1, a1b, f"{res} depth {depth} batch1 vs batch1 again")
        diff(a2, a2b, f"{res} depth {depth} batch2 vs batch2 again")
res = (3840, 2160)
s1 = render(scene, camf, res, 32, 1, 1, 1)
s1b = render(scene, camf, res, 32, 1, 1, 2)
diff(s1, s1b, "4K sample 1 only, batch param 1 vs 2")
for depth in (1, 2, 4, 8):
    a2 = render(scene, camf, res, depth, 1, 2, 2)
    a1 = render(scene, camf, res, depth, 1, 2, 1)
    diff(a2, a1, f"4K depth {depth} batch2 vs batch1")
