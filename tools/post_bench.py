"""Timing of the steps after the path on the bench scene at 1920x1080: fill_aux_buffers!, denoise!, postprocess!.
Host wall-clock around the C-ABI calls (so denoise / postprocess include their device->host copy of the result); the per-kernel
device times come from the ncu launch list taken over the same script (tools/gpu_post.sh).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hikari_jl_b200 import host as H, scenes          # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    scene, camf = scenes.c2_cat(256, 64)
    film = H.Film((1920, 1080))
    vp = H.VolPath(samples=4, max_depth=12)
    vp(scene, film, camf(film))
    n = 1920 * 1080
    out = {"workload": "C2 cat scene, 1920x1080, 4 spp accumulated", "pixels": n, "reps": reps}
    H.fill_aux_buffers(film, vp)                      # warm
    t = time.perf_counter()
    for _ in range(reps):
        vp.backend.call("fill_aux_buffers", 0)        # async launches ...
    vp.backend.call("read_aux_buffers", None, None, None)      # ... one stream sync
    out["fill_aux_ms"] = (time.perf_counter() - t) / reps * 1e3
    t = time.perf_counter()
    for _ in range(reps):
        H.fill_aux_buffers(film, vp)
    out["fill_aux_with_readback_ms"] = (time.perf_counter() - t) / reps * 1e3
    for name, cfg in (("denoise_default_ms", H.DenoiseConfig()), ("denoise_1pass_novar_ms", H.DenoiseConfig(iterations=1, use_variance=False))):
        H.denoise(film, vp, cfg)
        t = time.perf_counter()
        for _ in range(reps):
            H.denoise(film, vp, cfg)
        out[name] = (time.perf_counter() - t) / reps * 1e3
    for name, kw in (("postprocess_aces_ms", dict(tonemap="aces")), ("postprocess_masked_ms", dict(tonemap="aces", background=(0, 0, 0)))):
        H.postprocess(film, vp, **kw)
        t = time.perf_counter()
        for _ in range(reps):
            H.postprocess(film, vp, **kw)
        out[name] = (time.perf_counter() - t) / reps * 1e3
    out["hit_fraction"] = float((film.albedo[..., 0] > 0).mean())
    vp.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
