#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from an ncu report (needs -lineinfo).
   python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top] [FILE:LO-HI ...]   (NCU_SKIP=k picks the k-th matching launch; NCU_SORT=samples sorts by stall samples)"""
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
tot = {}
total_inst = 0
total_samp = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        inst = int(r[ii]); samp = int(r[si])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:110])
    a = tot.setdefault(key, [0, 0, {}])
    a[0] += inst; a[1] += samp
    for i, nm in stall_cols:
        try:
            a[2][nm] = a[2].get(nm, 0) + int(r[i])
        except ValueError:
            pass
    total_inst += inst; total_samp += samp
print(f"total warp-instructions {total_inst}, samples {total_samp}")
by = 1 if os.environ.get("NCU_SORT", "inst") == "samples" else 0
for (f, ln, src), (inst, samp, st) in sorted(tot.items(), key=lambda kv: -kv[1][by])[:top]:
    why = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v > 0)
    print(f"{100*inst/max(total_inst,1):5.1f}% inst {100*samp/max(total_samp,1):5.1f}% smp  {f}:{ln}  {src[:70]}  [{why}]")
allst = {}
for v in tot.values():
    for k, c in v[2].items():
        allst[k] = allst.get(k, 0) + c
print("stall samples:", " ".join(f"{k}={100*c/max(total_samp,1):.1f}%" for k, c in sorted(allst.items(), key=lambda kv: -kv[1]) if c))
# optional: totals per file / line range given as extra args FILE:LO-HI
for spec in sys.argv[4:]:
    f, rng = spec.split(":")
    lo, hi = map(int, rng.split("-"))
    t = sum(v[0] for (ff, ln, _), v in tot.items() if ff == f and lo <= ln <= hi)
    sm = sum(v[1] for (ff, ln, _), v in tot.items() if ff == f and lo <= ln <= hi)
    print(f"{spec}: {100*t/max(total_inst,1):.1f}% inst, {100*sm/max(total_samp,1):.1f}% samples")
