"""Small renders for compute-sanitizer (memcheck / racecheck / initcheck): a surface scene, a homogeneous-medium scene and a
NanoVDB cloud, each a few thousand paths so that the instrumented run stays in minutes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hikari_jl_b200 import host as H, scenes

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {
    "surfaces": (lambda: scenes.c1_spheres(8), (32, 24), 2, 4),
    "smoke": (scenes.cornell_smoke, (24, 24), 2, 4),
    "cloud": (lambda: scenes.c4_cloud((16, 16, 8), "nanovdb", (4, 4, 4)), (32, 18), 2, 6),
    "lights": (lambda: scenes.c3_many_lights(200, 8), (32, 18), 2, 4),
    "instanced": (lambda: scenes.c5_instanced(6, 8, instanced=True), (32, 18), 2, 4),
}
for name, (make, res, spp, depth) in cases.items():
    if which not in ("all", name):
        continue
    scene, camf = make()
    film = H.Film(res)
    vp = H.VolPath(samples=spp, max_depth=depth)
    img = vp(scene, film, camf(film)).copy()
    vp.close()
    print(name, "mean", float(img.mean()), "finite", bool(np.isfinite(img).all()), flush=True)
