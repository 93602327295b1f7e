#!/usr/bin/env python3
"""SASS instruction count per source line / file / function of one kernel (its .text section incl. the __noinline__ callees):
   python tools/sass_by_line.py obj.o 'kernel-mangled-name' [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, fun = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, by_line, by_file, by_fn = "?", collections.Counter(), collections.Counter(), collections.Counter()
inside, fn = False, "?"
for line in out.splitlines():
    if line.startswith("//---------------------"):
        inside = (".text." + fun) in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^(\$?[A-Za-z_.$][\w.$]*):", line.strip())
    if m and not line.strip().startswith(".L_"):
        fn = m.group(1)
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        by_line[cur] += 1; by_file[cur[0] if isinstance(cur, tuple) else cur] += 1; by_fn[fn] += 1
total = sum(by_line.values())
print("total", total)
print("by file:", dict(by_file.most_common(12)))
print("by function symbol:")
for k, v in by_fn.most_common(20): print(f"{v:7d}  {k[:110]}")
for (k, v) in by_line.most_common(top):
    print(f"{v:6d} {100.0 * v / total:5.1f}%  {k}")
