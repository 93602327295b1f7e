"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample-index partition + one all-reduce of the film
accumulators.  The renderer behind each rank is the CPU oracle; on the GPU box bench.py runs the same
logic over NCCL with libhikari_cuda.so."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hikari_jl_b200 import host as H, scenes
    import oracle_backend
    oracle_backend.lib().ok_set_num_threads(2)
    scene, camf = scenes.c1_spheres(8)
    res = (32, 24); spp = 4
    film = H.Film(res)
    vp = H.VolPath(samples=spp, max_depth=3, backend=oracle_backend.make_backend())
    vp._prepare(scene, film, camf(film)); vp.clear()
    vp.backend.call("render_samples_strided", rank + 1, world, spp // world)
    rgb, w = vp.backend.read_accum()
    acc = torch.from_numpy(np.concatenate([rgb.reshape(-1), w]))
    dist.all_reduce(acc)                                   # the single film reduce
    if rank == 0:
        n = res[0] * res[1]
        a = acc.numpy()
        img = (a[:3 * n].reshape(n, 3) / np.maximum(a[3 * n:], 1e-30)[:, None]).reshape(res[1], res[0], 3)
        np.save(os.path.join(out_dir, "dist.npy"), img)
        vp.clear(); film.iteration_index = 0
        vp.render(scene, film, camf(film), count=spp)
        np.save(os.path.join(out_dir, "single.npy"), film.framebuffer.copy())
    vp.close()
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_partition_and_film_reduce(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "dist.npy"); b = np.load(tmp_path / "single.npy")
    assert a.shape == b.shape and np.isfinite(a).all() and a.max() > 0
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)
