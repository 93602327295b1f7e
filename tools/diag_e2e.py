"""Where does the end-to-end frame loop lose time?  (C3, 4K, one sample per render! call.)
A: one batched call, ONE sample in flight (no lanes)       B: count=1 calls, no read-out (lanes overlap)
C: B + blocking read_film each frame                        D: B + async read-out (bench.py's e2e loop)
E: like D with the film finalize only (read into a device buffer)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN
from hikari_jl_b200 import host as H
import torch
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
BN.select_config(cfg)
scene, cam_of = BN.build_scene()
film = H.Film(BN.RES)
n = BN.RES[0] * BN.RES[1]
K = 24
DEPTH = int(os.environ.get('HK_E2E_DEPTH', '3'))
def run(tag, fn, batch):
    vp = H.VolPath(samples=4096, max_depth=BN.MAX_DEPTH, backend=H.Backend(device=0), sample_batch=batch)
    cam = cam_of(film)
    vp._prepare(scene, film, cam); vp.clear(); film.iteration_index = 0
    fn(vp, cam, 6); vp.backend.call("synchronize")
    t0 = time.perf_counter(); fn(vp, cam, K); vp.backend.call("synchronize"); dt = time.perf_counter() - t0
    print(f"{cfg} {tag}: {dt / K * 1e3:.3f} ms/frame  {n * K / dt / 1e6:.1f} Msamples/s", flush=True)
    vp.close()
def batched(vp, cam, k):
    vp.backend.call("render_samples", film.iteration_index + 1, k); film.iteration_index += k
def calls(vp, cam, k):
    for _ in range(k): vp.render(scene, film, cam, count=1, read=False)
def calls_noprep(vp, cam, k):
    for _ in range(k):
        vp.backend.call("render_samples", film.iteration_index + 1, 1); film.iteration_index += 1
def blocking(vp, cam, k):
    for _ in range(k): vp.render(scene, film, cam, count=1, read=True)
def asyncr(vp, cam, k):
    pend = []
    for _ in range(k):
        pend.append(vp.render(scene, film, cam, count=1, read="async"))
        if len(pend) > DEPTH: vp.wait_film(film, pend.pop(0))
    while pend: vp.wait_film(film, pend.pop(0))
if os.environ.get("HK_LANE_MAX_COUNT"):
    M = int(os.environ["HK_LANE_MAX_COUNT"])
    def chunks(vp, cam, k):
        for _ in range(k // M):
            vp.backend.call("render_samples", film.iteration_index + 1, M); film.iteration_index += M
    K = 32
    run(f"L lanes x {M} samples per call", chunks, M)
    run("A16 batched, 16 in flight", batched, 16)
    sys.exit(0)
run("A batched, 1 sample in flight", batched, 1)
run("A2 batched, 2 samples in flight", batched, 2)
run("A3 batched, 3 samples in flight", batched, 3)
run("A16 batched, 16 in flight", batched, 16)
run("B count=1 calls, no read", calls, 1)
run("B2 count=1 calls, no set_camera", calls_noprep, 1)
run("C blocking read", blocking, 1)
run("D async read", asyncr, 1)
