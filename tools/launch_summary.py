#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total time and launch count per kernel name, share of the total."""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(u, 1e-6)
    name = re.sub(r"\(.*", "", r[ik]); tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.3f} ms over {sum(cnt.values())} launches")
for k, v in tot.most_common(30): print(f"{v:10.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:5d}  {k}")
