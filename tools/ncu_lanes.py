#!/usr/bin/env python
"""Per-source-line warp-instruction share and active lanes per instruction from an ncu report (needs -lineinfo).
   python tools/ncu_lanes.py REPORT.ncu-rep KERNEL_REGEX [top]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = None
tot = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        a, b = int(r[ii]), int(r[ti])
    except ValueError:
        continue
    t = tot.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:100]])
    t[0] += a
    t[1] += b
T = sum(v[0] for v in tot.values())
byfile, byfile_t = collections.Counter(), collections.Counter()
for (f, l), v in tot.items():
    byfile[f] += v[0]
    byfile_t[f] += v[1]
print(f"total warp-instructions {T}")
for f in byfile:
    print(f"  {f}: {100 * byfile[f] / T:.1f}% of instructions, {byfile_t[f] / max(1, byfile[f]):.1f} lanes")
for (f, l), v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / T:5.1f}% lanes {v[1] / max(1, v[0]):4.1f} {f}:{l} {v[2]}")
