#!/usr/bin/env python
"""Per-source-line instruction counts of an .ncu-rep captured with --import-source on (kernel built with -lineinfo):
usage: ncu_lines.py REPORT.ncu-rep [TOP]  -> lines of csrc/*.cuh|cu sorted by warp instructions executed, with lane utilisation."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
total = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ia = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0].isdigit() and r[ia].isdigit():      # a CUDA source line (aggregated over its SASS)
        n, t = int(r[ia] or 0), int(r[it] or 0)
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += n; a[1] += t; total += n
print("total warp instructions attributed:", total)
# function-level buckets by line ranges are easier to read: print top lines
for (f, ln), (n, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:12d} {100.0*n/total:5.1f}%  lanes {t/max(n,1):5.1f}  {f}:{ln}  {src}")
