#!/usr/bin/env python3
"""bench.py — VolPath hot-path benchmark (BASELINE.json metric: Msamples/s & Mrays/s, fraction of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # libhikari_cuda.so on N B200s (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path restated (oracle), host cores

A "step" is one render!() pass = one sample per pixel over the whole frame (src/integrators/volpath/volpath.jl:445-636).
Workload = BASELINE.json configs[2] restated: glass sphere + gold plane under a sun/sky environment light and
10 000 emissive triangles (light BVH), 3840x2160, max_depth 12 -- one of the 4K configurations the targets are stated on; an
N = 1 run also appends a `per_config` table (C1..C5, each at its real resolution).  The CPU legs render the same frame.
Multi-GPU: scene replicated, sample indices partitioned round-robin over ranks (rank g renders g+1, g+1+N, ...),
one NCCL all-reduce of the film accumulators (4*W*H f32) at the end, inside the timed region.
"""
import argparse
import ctypes as C
import json
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before anything creates the CUDA context (libhikari_cuda.so does the same on load;
                                                                 # hk_api.cu: the render lanes need more than the default 8 hardware work queues)
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY 8d configs restated.  The default (the configuration the driver's BENCH / SCALE lines are quoted on) is C3, one of the
# 4K configurations the north_star's targets name; every run at N = 1 also carries a `per_config` table over C1..C5.
# cpu_rows: the CPU legs (cpu_baseline, --impl reference) render the SAME scene, camera and resolution, but only every
# cpu_rows-th image row per step (a bounded sample of the frame: ~0.5 M camera rays, 1-2 s on 16 threads).
CONFIGS = {
    "C1": dict(workload="C1 sphere-normals scene, 512x512, VolPath max_depth=5, 1 spp per step", res=(512, 512), depth=5, cpu_rows=1),
    "C2": dict(workload="C2 cat scene (procedural stand-in mesh), 1920x1080, VolPath max_depth=12, 1 spp per step", res=(1920, 1080), depth=12, cpu_rows=4),
    "C3": dict(workload="C3 glass sphere + gold plane, sun/sky env light + 10k emissive triangles (light BVH), 3840x2160, max_depth=12, 1 spp per step", res=(3840, 2160), depth=12, cpu_rows=16),
    "C4": dict(workload="C4 procedural cumulus NanoVDB cloud (256x256x128), delta tracking, 3840x2160, max_depth=32, 1 spp per step", res=(3840, 2160), depth=32, cpu_rows=32),
    "C5": dict(workload="C5 instanced blob meshes (mixed materials), sun/sky env light, 3840x2160, max_depth=8, 1 spp per step", res=(3840, 2160), depth=8, cpu_rows=16),
}
DEFAULT_CONFIG = "C3"
WORKLOAD, RES, MAX_DEPTH, CPU_ROWS = None, None, None, None
C5_INSTANCES = 1000


def select_config(name, c5_instances=1000):
    global WORKLOAD, RES, MAX_DEPTH, CPU_ROWS, CONFIG_NAME, C5_INSTANCES
    c = CONFIGS[name]
    CONFIG_NAME, WORKLOAD, RES, MAX_DEPTH, CPU_ROWS, C5_INSTANCES = name, c["workload"], c["res"], c["depth"], c["cpu_rows"], c5_instances
    if name == "C5":
        WORKLOAD += f" ({c5_instances} instances)"


select_config(DEFAULT_CONFIG)


def config_dict(scene, world):
    """`config` of the JSON line: identical for the CUDA arm and the --impl reference arm of the same command."""
    n = RES[0] * RES[1]
    return {"workload": WORKLOAD, "name": CONFIG_NAME, "resolution": [RES[0], RES[1]], "max_depth": MAX_DEPTH,
            "triangles": int(scene.triangle_count()), "lights": int(len(scene._synced.lights)),
            "partition": f"sample-index round-robin x{world}",
            "l2_note": f"per-pass working set (path state + queues, ~{288 * n / 1e9:.2f} GB per sample in flight) exceeds the 126 MB L2; no explicit flush"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks + throttle reasons while the timed region runs: NVML in-process every ~2 ms (the timed region of
    the default run is only tens of milliseconds), nvidia-smi as the fallback when pynvml is not importable."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0])); self.max_mhz = float(out[1])
        for nm, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.002 if self.nvml is not None else 0.2)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "n_samples": len(self.samples),
                "how": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_scene():
    from hikari_jl_b200 import scenes
    if CONFIG_NAME == "C1":
        return scenes.c1_spheres(64)
    if CONFIG_NAME == "C3":
        return scenes.c3_many_lights(10000, 128)
    if CONFIG_NAME == "C4":
        return scenes.c4_cloud((256, 256, 128), "nanovdb", (64, 64, 64))
    if CONFIG_NAME == "C5":
        return scenes.c5_instanced(C5_INSTANCES, 160, instanced=True)
    return scenes.c2_cat(256, 64)


def _oracle_vp():
    """The reference's CPU implementation of the path, restated (oracle/): Julia + Raycore.jl are not installable in this image,
    so both CPU legs time the C++ port (kind "port") on the SAME scene, camera, resolution and max_depth as the CUDA arm; a step
    renders every CPU_ROWS-th image row of the frame (ok_set_row_subset), the row offset rotating from step to step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build_oracle()
    import oracle_backend
    from hikari_jl_b200.host import Film, VolPath
    scene, camf = build_scene()
    film = Film(RES)
    vp = VolPath(samples=4096, max_depth=MAX_DEPTH, backend=oracle_backend.make_backend())
    vp._prepare(scene, film, camf(film)); vp.clear()
    olib = oracle_backend.lib()
    if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        olib.ok_set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP to 1 thread; rank 0 runs alone: use the box
    rows = (RES[1] + CPU_ROWS - 1) // CPU_ROWS

    def step(k):
        olib.ok_set_row_subset(vp.backend.ctx, CPU_ROWS, k % CPU_ROWS)
        vp.backend.call("render_samples", k + 1, 1)
        return RES[0] * len(range(k % CPU_ROWS, RES[1], CPU_ROWS))
    sample = (f"same scene / camera / {RES[0]}x{RES[1]} / max_depth {MAX_DEPTH}; each step renders 1 spp of every {CPU_ROWS}th image row "
              f"(~{RES[0] * rows} camera rays per step)") if CPU_ROWS > 1 else f"same scene / camera / {RES[0]}x{RES[1]} / max_depth {MAX_DEPTH}; each step renders 1 spp of the whole frame"
    return scene, vp, olib, step, sample


def run_reference(args):
    """--impl reference: times the CPU arm with all host threads; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scene, vp, olib, step, sample = _oracle_vp()
    cores = olib.ok_num_threads()
    for w in range(args.warmup):
        step(w)
    r0 = olib.ok_rays_traced(vp.backend.ctx)
    t0 = time.perf_counter()
    pixels = 0
    for k in range(args.steps):
        pixels += step(args.warmup + k)
    dt = time.perf_counter() - t0
    rays = olib.ok_rays_traced(vp.backend.ctx) - r0
    val = pixels / dt / 1e6
    line = {
        "impl": "reference", "metric": "VolPath throughput", "value": val, "unit": "Msamples/s", "n_gpus": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "mrays_per_s": rays / dt / 1e6,
        "config": config_dict(scene, world), "sample": sample,
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    vp.close()
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(budget_s=12.0):
    scene, vp, olib, step, sample = _oracle_vp()
    step(0)
    t0 = time.perf_counter(); steps = 0; pixels = 0
    while steps < 2 or (time.perf_counter() - t0 < budget_s and steps < 16):
        pixels += step(1 + steps); steps += 1
    dt = time.perf_counter() - t0
    cores = olib.ok_num_threads()
    vp.close()
    return {"value": pixels / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample + f"; {steps} steps"}


class _DevPtr:
    """Expose a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


STAGES = ["camera", "trace", "medium", "escaped", "shade", "shadow", "film", "route"]


def measure(args, rank, world, local, dist, torch, full):
    """One configuration (the one select_config() chose) on this rank's GPU.  Returns the JSON-line dict on rank 0 (None elsewhere).
    full: the headline configuration (e2e loop, CPU baseline, collective); otherwise a short per_config row."""
    from hikari_jl_b200 import _abi as A
    from hikari_jl_b200.host import Backend, Film, VolPath
    scene, camf = build_scene()
    film = Film(RES)
    vp = VolPath(samples=4096, max_depth=MAX_DEPTH, backend=Backend(device=local), sample_batch=args.batch)
    cam = camf(film)
    t_up0 = time.perf_counter()
    vp._prepare(scene, film, cam)
    t_upload = time.perf_counter() - t_up0
    vp.clear()
    B, lib, ctx = vp.backend, vp.backend.lib, vp.backend.ctx
    n = RES[0] * RES[1]
    steps = args.steps
    stats = A.HkStats()
    accp, accn = C.c_void_p(), C.c_uint64()
    lib.hk_film_accum_dev(ctx, C.byref(accp), C.byref(accn))
    acc_t = torch.as_tensor(_DevPtr(accp.value, accn.value), device=f"cuda:{local}")

    def sample_of(k):      # rank-strided sample indices: the multi-GPU partition
        return rank + 1 + k * world

    # W untimed steps in one call like the timed region; if W is smaller than the number of samples the timed call keeps in
    # flight, the warm-up is topped up to that number so that the path-state pool (allocated on demand) has its final size
    # and has been touched before the timed region starts.  The film is cleared afterwards: the timed region renders sample
    # indices the warm-up has not touched, so the film that is all-reduced holds exactly N x K distinct samples.
    auto_batch = max(1, min(64, (128 << 20) // n))                      # HK_AUTO_SLOTS in hk_api.cu
    batch_used = min(args.batch if args.batch > 0 else auto_batch, steps)
    warm = max(args.warmup, batch_used)
    B.call("render_samples_strided", sample_of(0), world, warm)
    vp.clear()
    if dist and full:        # warm the collective too (NCCL sets up its channels on the first call of a given size); on a scratch
        scratch = torch.zeros_like(acc_t)                            # buffer: the film accumulators are reduced exactly once
        for _ in range(2):
            dist.all_reduce(scratch)
        del scratch
    B.call("synchronize")
    # ---- timed region: K steps, device-timed (CUDA events on the library's launch stream), barrier + sync both sides ----
    if dist: dist.barrier()
    torch.cuda.synchronize()
    lib.hk_stats(ctx, C.byref(stats)); rays0, launches0, verts0 = stats.rays_traced, stats.kernel_launches, stats.path_vertices
    sampler = ClockSampler(local); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    B.call("render_samples_strided", sample_of(warm), world, steps)      # K steps in ONE call: no host sync inside
    B.call("synchronize")
    lib.hk_stats(ctx, C.byref(stats))
    dev_ms = float(stats.last_render_ms)                                               # events recorded on the launching stream
    red_ms = 0.0
    if dist:
        ev0.record(); dist.all_reduce(acc_t); ev1.record(); torch.cuda.synchronize(); red_ms = ev0.elapsed_time(ev1)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.result()
    rays = stats.rays_traced - rays0
    verts = stats.path_vertices - verts0
    launches = stats.kernel_launches - launches0
    t_ms = dev_ms + red_ms
    if dist:
        tt = torch.tensor([t_ms, float(rays)], device=f"cuda:{local}", dtype=torch.float64)
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_ms, rays = float(tmax[0]), int(tsum[1])
    value = world * n * steps / (t_ms * 1e-3) / 1e6
    e2e = None
    if full:
        # ---- e2e: the interactive render! loop through the public API with HOST buffers: per step the camera is re-sent
        # (H2D), one sample pass runs and the framebuffer is read back (D2H) ------------------------------------------------
        e2e_steps = max(3, steps)                      # the same K frames as the device-timed region
        vp.clear(); film.iteration_index = 0
        if dist: dist.barrier()
        depth = max(1, min(7, int(os.environ.get("HK_E2E_DEPTH", "3"))))      # read-outs left in flight while the next frame is enqueued
        def display_loop(frames):        # progressive display loop: frame k's read-out (own stream) and frame k+1's render overlap frame k+2's
            pending = []
            for _ in range(frames):
                pending.append(vp.render(scene, film, cam, count=1, read="async"))
                if len(pending) > depth:
                    vp.wait_film(film, pending.pop(0))
            while pending:
                vp.wait_film(film, pending.pop(0))      # every frame has landed in host memory before the loop returns
        display_loop(6)                  # untimed, in the SAME pipelined pattern as the timed loop: the page-locked host buffers (displayed + in flight),
        vp.clear(); film.iteration_index = 0      # the copy stream and the staging buffers all come into being here, not in the timed region
        B.call("synchronize")
        if dist: dist.barrier()
        t0 = time.perf_counter()
        display_loop(e2e_steps)
        e2e_dt = time.perf_counter() - t0
        if dist:
            te = torch.tensor([e2e_dt], device=f"cuda:{local}", dtype=torch.float64); dist.all_reduce(te, op=dist.ReduceOp.MAX); e2e_dt = float(te[0])
        e2e = {"value": world * n * e2e_steps / e2e_dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": C.sizeof(A.HkCamera), "d2h_bytes_per_step": 12 * n,
               "steps": e2e_steps, "what": "per step: camera H2D, render!(vp, scene, film, camera) of one sample, framebuffer D2H into page-locked host memory (pipelined: hk_read_film_async / _wait, " + str(depth) + " read-outs left in flight while the next frame is enqueued)",
               "scene_upload_s": t_upload}
    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (k_trace): per-launch device time + traversal work counters, measured live, over
        # the passes of fresh sample indices (the film is not read again) -------------------------------------------------
        lib.hk_set_profiling(ctx, 1)
        B.call("render_samples_strided", sample_of(warm + steps), world, steps)
        ms = (C.c_double * 8)(); ln = (C.c_uint64 * 8)(); wk = (C.c_uint64 * 6)()
        lib.hk_stage_times(ctx, ms, ln, wk)
        stage_ms = list(ms); stage_ln = list(ln)
        lib.hk_set_profiling(ctx, 2)
        B.call("render_samples_strided", sample_of(warm + steps), world, steps)
        lib.hk_stage_times(ctx, ms, ln, wk)
        lib.hk_set_profiling(ctx, 0)
        work = list(wk)
        peak, peak_src = load_peaks()
        trace_bytes = work[0] * 48 + work[1] * 80 + work[2] * 48          # 32 B ray + 16 B hit + 80 B/node + 48 B/tri
        trace_ms = stage_ms[1]
        achieved = trace_bytes / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        total_ms = sum(stage_ms) or 1.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", f"trace_traffic_{CONFIG_NAME}.json")     # dram__bytes_read+write per launch, one ncu --set full capture
        if os.path.exists(tpath):
            tj = json.load(open(tpath)); traffic = tj["dram_bytes_per_launch"]; traffic_src = tj["source"]
        roofline = {"kernel": "k_trace (closest-hit BVH8 traversal)", "bound": "hbm", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "bytes_per_launch": trace_bytes / max(1, stage_ln[1]), "ms_per_launch": trace_ms / max(1, stage_ln[1]),
                    "rays": work[0], "node_visits_per_ray": work[1] / max(1, work[0]), "tri_tests_per_ray": work[2] / max(1, work[0]),
                    "stage_share": {nm: stage_ms[i] / total_ms for i, nm in enumerate(STAGES)},
                    "shadow": {"rays": work[3], "node_visits_per_ray": work[4] / max(1, work[3]), "tri_tests_per_ray": work[5] / max(1, work[3]),
                               "achieved": (work[3] * 48 + work[4] * 80 + work[5] * 48) / max(1e-9, stage_ms[5] * 1e-3) / 1e9}}
        # whole path, SURVEY 8d: per path vertex the reference moves 2 178 B of queue records; + the traversal bytes of every
        # closest-hit and shadow query (counted by the COUNT pass over the same number of samples) + 84 B per pixel sample of film traffic
        whole_bytes = verts * 2178 + trace_bytes + (work[3] * 48 + work[4] * 80 + work[5] * 48) + n * steps * 84
        whole_achieved = whole_bytes / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else 0.0
        roofline["whole_path"] = {"algorithmic_bytes_per_step": whole_bytes / steps, "path_vertices_per_sample": verts / (n * steps),
                                  "achieved": whole_achieved, "peak": peak, "unit": "GB/s", "frac": whole_achieved / peak,
                                  "note": "reference AOS record traffic (SURVEY 8d) for the work done, divided by this rank's device time"}
        roofline["stage_ms_per_step"] = {nm: stage_ms[i] / steps for i, nm in enumerate(STAGES)}
        cpu = cpu_baseline_leg() if (full and not args.quick and world == 1) else None       # the CPU baseline is reported at N=1 only
        line = {
            "metric": "VolPath throughput", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": rays / (t_ms * 1e-3) / 1e6, "rays_per_sample": rays / (world * n * steps),
            "config": config_dict(scene, world), "samples_in_flight": int(batch_used), "warmup_steps_run": int(warm),
            "e2e": e2e, "gpu_launches": int(launches), "wall_s": wall, "film_reduce_ms": red_ms, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
    vp.close()
    return line


def run_cuda(args):
    # torchrun exports OMP_NUM_THREADS=1; the host-side scene build (BVH build, OpenMP) would then run on one
    # core per rank.  Give every rank its share of the cores instead (must happen before the OpenMP runtime is loaded).
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // int(os.environ["WORLD_SIZE"])))
    import torch
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libhikari_cuda.so has no CPU fallback")
    if rank == 0:
        g.build_cuda()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    line = measure(args, rank, world, local, dist, torch, full=True)
    # ---- per_config: every SURVEY 8d configuration at its real resolution on one GPU (short runs; N = 1 only) --------------
    if world == 1 and not args.no_per_config and line is not None:
        table = {}
        main_name = CONFIG_NAME
        for name in sorted(CONFIGS):
            if name == main_name:
                row = line
            else:
                select_config(name, args.c5_instances)
                sub = argparse.Namespace(**vars(args)); sub.steps = min(args.steps, 16); sub.quick = True
                try:
                    row = measure(sub, 0, 1, local, None, torch, full=False)
                except Exception as e:       # a config that cannot run (e.g. out of memory on a shared box) must not take the line down
                    table[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
                    continue
            rf = row["roofline"]
            table[name] = {"workload": row["config"]["workload"], "value": row["value"], "unit": "Msamples/s", "ms_per_step": row["ms_per_step"],
                           "mrays_per_s": row["mrays_per_s"], "steps": row["steps"], "samples_in_flight": row["samples_in_flight"],
                           "triangles": row["config"]["triangles"], "k_trace_frac": rf["frac"], "k_trace_gbs": rf["achieved"],
                           "whole_path_frac": rf["whole_path"]["frac"], "stage_ms_per_step": rf["stage_ms_per_step"], "gpu_launches": row["gpu_launches"]}
        select_config(main_name, args.c5_instances)
        line["per_config"] = table
    if dist:
        dist.barrier(); dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="samples kept in flight per wavefront pass (HkRenderParams.sample_batch); 0 = the library's automatic choice")
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS), help="SURVEY 8d workload (default C3 = BASELINE.json configs[2], a 4K configuration)")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config table (C1..C5 short runs) that an N = 1 run appends")
    ap.add_argument("--c5-instances", type=int, default=1000, help="C5 only: instances of the ~50k-triangle base mesh (1000 = 50 M triangles)")
    ap.add_argument("--quick", action="store_true", help="development: skip the CPU baseline leg (tuning-variant sweeps, tools/variants.py)")
    args = ap.parse_args()
    select_config(args.config, args.c5_instances)
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
