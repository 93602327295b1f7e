#!/usr/bin/env python3
"""Tuning-variant sweeps of libhikari_cuda.so (development tool).

  python tools/variants.py build name1:-DHK_X=1,-DHK_Y=2 name2:...   # here (nvcc cross-compiles), into build/variants/
  python tools/variants.py run [--steps K]                          # on the GPU box (through gpurun): bench.py --quick per variant
  python tools/variants.py table                                    # here: summarise gpurun_out/var_*.json
"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")


def build(specs):
    import __graft_entry__ as g
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "*.so*")): os.remove(f)
    for spec in specs:
        name, defs, envs = (spec.split(":") + ["", ""])[:3]          # name:-DA=1,-DB=2:ENV1=x,ENV2=y
        out = os.path.join(VDIR, f"lib_{name}.so")
        open(out + ".env", "w").write(envs)
        try:
            g.build_cuda(extra_flags=tuple(d for d in defs.split(",") if d), out=out)
            for d in glob.glob(os.path.join(ROOT, "build", "obj_*")):      # the variant's objects are not needed once it is linked (and gpurun snapshots are size-limited)
                import shutil; shutil.rmtree(d, ignore_errors=True)
            print(name, "OK")
        except Exception as e:
            print(name, "FAILED", e)


def run(argv):
    steps = argv[argv.index("--steps") + 1] if "--steps" in argv else "8"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for so in sorted(glob.glob(os.path.join(VDIR, "*.so"))):
        name = os.path.basename(so)[4:-3]
        env = dict(os.environ, HK_CUDA_LIB=so)
        if os.path.exists(so + ".env"):
            env.update(dict(kv.split("=", 1) for kv in open(so + ".env").read().split(",") if "=" in kv))
        for cfg in os.environ.get("HK_BENCH_CONFIGS", "C2").split(","):          # e.g. HK_BENCH_CONFIGS="C2,C3,C4"
            extra = os.environ.get("HK_BENCH_ARGS", "").split()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--no-per-config", "--config", cfg, "--steps", steps, "--warmup", "3"] + extra,
                               env=env, capture_output=True, text=True, timeout=900)
            open(os.path.join(ROOT, "gpurun_out", f"var_{name}_{cfg}.json"), "w").write(r.stdout if r.returncode == 0 else json.dumps({"error": r.stderr[-1500:]}))
            print(name, cfg, r.stdout[:160] if r.returncode == 0 else r.stderr[-500:], flush=True)


def table():
    for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "var_*.json"))):
        try: j = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception as e: print(os.path.basename(f), "unreadable", e); continue
        if "error" in j: print(os.path.basename(f), "ERROR", j["error"][-300:]); continue
        st = j["roofline"].get("stage_ms_per_step", {})
        print(f"{os.path.basename(f)[4:-5]:28s} {j['value']:7.1f} Msamples/s  {j['ms_per_step']:6.3f} ms/step  " + "  ".join(f"{k}={v:.3f}" for k, v in st.items()) + f"  frac={j['roofline']['frac']:.3f}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "build": build(sys.argv[2:])
    elif cmd == "run": run(sys.argv[2:])
    else: table()
