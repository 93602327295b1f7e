#!/usr/bin/env python3
"""dram__bytes_read/write per launch of the captured k_trace launches of an .ncu-rep -> profiles/trace_traffic_<config>.json
   python tools/trace_traffic.py rep.ncu-rep C3 "<the command that was profiled>" """
import csv, io, json, subprocess, sys
rep, cfg, src = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[ix[k]].replace(",", "")); u = units[ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1)
rd = [val(r, "dram__bytes_read.sum") for r in rows[2:]]; wr = [val(r, "dram__bytes_write.sum") for r in rows[2:]]
tm = [val(r, "gpu__time_duration.sum") for r in rows[2:]]
import math
keep = [i for i in range(len(rd)) if not (math.isnan(rd[i]) or math.isnan(wr[i]))]      # (ncu occasionally returns no counters for a launch)
rd, wr, tm = [rd[i] for i in keep], [wr[i] for i in keep], [tm[i] for i in keep]
n = len(rd)
out = {"source": src, "kernel": rows[2][ix["Kernel Name"]] if n else None, "launches": n,
       "dram_bytes_read_per_launch": sum(rd) / max(n, 1), "dram_bytes_write_per_launch": sum(wr) / max(n, 1),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / max(n, 1), "per_launch_dram_bytes": [a + b for a, b in zip(rd, wr)], "per_launch_time_s_under_ncu": tm}
json.dump(out, open(f"gpurun_out/trace_traffic_{cfg}.json", "w"), indent=1)
print(json.dumps({k: out[k] for k in ("launches", "dram_bytes_per_launch")}))
