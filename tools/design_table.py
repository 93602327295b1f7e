#!/usr/bin/env python3
"""Markdown tables for DESIGN.md / README.md from a bench JSON line: python tools/design_table.py gpurun_out/bench.json"""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j["roofline"]
print("| B200, round 2 (`%s`) | value |" % sys.argv[1].replace("gpurun_out", "profiles"))
print("|---|---:|")
print(f"| C3 device-only (`value`) | **{j['value']:.0f} Msamples/s**, {j['ms_per_step']:.2f} ms/step, {j['mrays_per_s']:.0f} Mrays/s, {j['samples_in_flight']} samples in flight |")
print(f"| C3 end to end (`e2e`: host buffers, H2D + D2H inside) | {j['e2e']['value']:.0f} Msamples/s |")
c = j.get("cpu_baseline")
if c: print(f"| CPU arm (oracle port, {c['cores']} threads, same config, row-subset sample) | {c['value']:.3f} Msamples/s |")
print(f"| k_trace roofline | {r['achieved']:.0f} GB/s algorithmic = **{r['frac']:.2f}** of {r['peak']:.0f} GB/s; DRAM traffic per launch {(r['traffic'] or 0) / 1e6:.0f} MB vs {r['bytes_per_launch'] / 1e6:.0f} MB algorithmic |")
print(f"| whole path (SURVEY 8d record traffic) | {r['whole_path']['achieved']:.0f} GB/s = {r['whole_path']['frac']:.2f} |")
print(f"| kernels launched in the timed region | {j['gpu_launches']} |")
print()
print("| config | triangles | samples in flight | Msamples/s | Mrays/s | ms/step | largest stages (ms/step) | k_trace frac | whole-path frac |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|")
for k, v in sorted(j.get("per_config", {}).items()):
    st = sorted(v["stage_ms_per_step"].items(), key=lambda kv: -kv[1])[:3]
    print(f"| {v['workload'].split(',')[0]} | {v['triangles']:,} | {v['samples_in_flight']} | {v['value']:.0f} | {v['mrays_per_s']:.0f} | {v['ms_per_step']:.2f} | " + ", ".join(f"{a} {b:.2f}" for a, b in st) + f" | {v['k_trace_frac']:.2f} | {v['whole_path_frac']:.2f} |")
