#!/usr/bin/env python3
"""Per-source-line hot spots of one captured launch: tools/ncu_hot_lines.py rep.ncu-rep [launch_skip] [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": print("kernel:", r[1]); continue
    if r[0] == "Line No": hdr = r; i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        key = (cur_file, r[0]); a = agg.setdefault(key, [r[1], 0, 0, 0])
        a[1] += int(r[i_s] or 0); a[2] += int(r[i_e] or 0); a[3] += int(r[i_t] or 0)
ts = sum(a[1] for a in agg.values()) or 1; te = sum(a[2] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {te}")
byfile = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[1]
print("samples by file:", ", ".join(f"{f}={v/ts*100:.1f}%" for f, v in byfile.most_common()))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{a[1]/ts*100:5.1f}% samp {a[2]/te*100:5.1f}% inst {a[3]/max(1,a[2]):5.1f} lanes  {f}:{l:>4}  {a[0].strip()[:100]}")
