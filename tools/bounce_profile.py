#!/usr/bin/env python3
"""Per-bounce queue counts and stage times of one sample pass of the bench workload (GPU; diagnostics)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hikari_jl_b200.host import Backend, Film, VolPath

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if len(sys.argv) > 2: bench.select_config(sys.argv[2])      # python tools/bounce_profile.py <batch> <config>
scene, camf = bench.build_scene()
film = Film(bench.RES)
vp = VolPath(samples=4096, max_depth=bench.MAX_DEPTH, sample_batch=batch, backend=Backend())
vp._prepare(scene, film, camf(film)); vp.clear()
B, lib, ctx = vp.backend, vp.backend.lib, vp.backend.ctx
B.call("render_samples", 1, 3 * batch)
lib.hk_set_profiling(ctx, 5)
B.call("render_samples", 1 + 3 * batch, batch)
D = bench.MAX_DEPTH
cnt = (C.c_uint32 * (16 * D))(); ms = (C.c_double * (8 * D))()
lib.hk_bounce_profile.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
lib.hk_bounce_profile(ctx, D, cnt, ms)
lib.hk_set_profiling(ctx, 0)
names = ["camera", "trace", "medium", "escaped", "shade", "shadow", "film", "route"]
print(f"batch={batch}  bounce  rays_in  escaped  shadow  hits | ms: " + " ".join(names[1:6] + names[7:]))
tot = 0.0
for d in range(D):
    c = cnt[16 * d:16 * d + 16]; m = ms[8 * d:8 * d + 8]
    rays_in = c[d & 1]; tot += sum(m)
    print(f"  {d:2d} {rays_in:9d} {c[2]:8d} {c[4]:8d} {c[5]:8d} | " + " ".join(f"{m[i]:7.3f}" for i in (1, 2, 3, 4, 5, 7)) + f"  sum={sum(m):.3f}")
print(f"total bounce ms {tot:.3f} per pass of {batch} sample(s) -> {tot / batch:.3f} ms/sample")
vp.close()
