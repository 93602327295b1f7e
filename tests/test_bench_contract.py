"""CPU test of the bench.py contract: the reference arm (the oracle port on the host cores) prints ONE JSON line with the
keys the driver reads, and the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from util import gpu_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "VolPath throughput" and j["unit"] == "Msamples/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "C1" in j["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_cuda_arm_fails_loudly_without_a_device():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_config_is_the_cuda_arm_config():
    """Both arms print config_dict(): same workload name, resolution, max_depth, triangle and light counts (the driver compares them)."""
    sys.path.insert(0, ROOT)
    import bench
    bench.select_config("C1")
    scene, _ = bench.build_scene()
    c = bench.config_dict(scene, 1)
    assert c["resolution"] == [512, 512] and c["max_depth"] == 5 and c["name"] == "C1" and c["triangles"] == 23826
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    j = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
    assert j["config"] == c
    assert bench.DEFAULT_CONFIG == "C3" and bench.CONFIGS["C3"]["res"] == (3840, 2160), "the driver's bench line is quoted on a 4K configuration"


def test_row_subset_is_a_sample_of_the_same_frame():
    """The CPU legs render every k-th image row of the full-resolution frame (ok_set_row_subset): those rows are bit-identical to
    the same rows of the full render (per-pixel sample streams are independent of the other pixels), the rest stay black."""
    import numpy as np
    import oracle_backend
    from hikari_jl_b200 import scenes
    from hikari_jl_b200.host import Film, VolPath
    scene, camf = scenes.c1_spheres(12)
    outs = []
    for step, off in ((1, 0), (4, 1)):
        film = Film((48, 32))
        vp = VolPath(samples=2, max_depth=4, backend=oracle_backend.make_backend())
        vp._prepare(scene, film, camf(film)); vp.clear()
        assert oracle_backend.lib().ok_set_row_subset(vp.backend.ctx, step, off) == 0
        vp.backend.call("render_samples", 1, 2)
        vp.backend.read_film(film.framebuffer)
        outs.append(film.framebuffer.copy()); vp.close()
    full, sub = outs
    # framebuffer[py, px] with py = H - y (the film is flipped vertically w.r.t. pixel rows y = 1..H, volpath.jl:384-417)
    rows = np.array([r for r in range(32) if r % 4 == 1])
    lit = np.zeros(32, bool)
    for axis_rows in (rows, 31 - rows):
        if np.array_equal(sub[axis_rows].view(np.uint32), full[axis_rows].view(np.uint32)) and sub[axis_rows].max() > 0:
            lit[axis_rows] = True
    assert lit.sum() == len(rows), "the sampled rows must equal the full render's rows bit for bit"
    assert (sub[~lit] == 0).all()
