"""Development: per-ray work statistics of the tracking kernels on C4 (needs a library built with -DHK_MEDIA_STATS, HK_CUDA_LIB=...)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hikari_jl_b200 import host as H, scenes, _abi as A
scene, camf = scenes.c4_cloud((256, 256, 128), "nanovdb", (64, 64, 64))
film = H.Film((1920, 1080)); vp = H.VolPath(samples=4096, max_depth=32, sample_batch=1)
vp._prepare(scene, film, camf(film)); vp.clear()
lib = A.load_library()
out = (C.c_ulonglong * 32)()
lib.hk_dev_media_stats(out)
vp.backend.call("render_samples", 1, 1); vp.backend.call("synchronize")
lib.hk_dev_media_stats(out)
v = list(out)
names = ["delta rays", "delta empty cells", "delta fetched cells", "delta events", "delta seg ends", "delta skip iters(warp)", "delta event iters(warp)", "delta event lane-iters",
         "ratio rays", "ratio empty cells", "ratio fetched cells", "ratio events", "ratio seg ends", "ratio event lane-iters", "ratio event iters(warp)", "delta skip lane-iters"]
for n, x in zip(names, v): print(f"{n:28s} {x:14d}")
print("per delta ray: empty %.1f fetched %.1f events %.2f seg-ends %.2f" % (v[1] / max(v[0], 1), v[2] / max(v[0], 1), v[3] / max(v[0], 1), v[4] / max(v[0], 1)))
print("per ratio ray: empty %.1f fetched %.1f events %.2f seg-ends %.2f" % (v[9] / max(v[8], 1), v[10] / max(v[8], 1), v[11] / max(v[8], 1), v[12] / max(v[8], 1)))
print("delta: lanes per event iteration %.1f, per skip iteration %.1f" % (v[7] / max(v[6], 1), v[15] / max(v[5], 1)))
print("delta rays with >= 1 segment %d, with none %d; ratio segments with >= 1 majorant segment %d, with none %d, vacuum %d" % (v[16], v[17], v[18], v[19], v[20]))
print("ratio: lanes per event iteration %.1f" % (v[13] / max(v[14], 1)))
