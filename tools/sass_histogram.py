#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libaugcuda.so (cuobjdump -sass), for profiles/.

usage: python tools/sass_histogram.py [path/to/libaugcuda.so] > profiles/sass_opcodes_rNN.txt

For every kernel: instruction count, registers are in the ncu summaries; here the opcode classes that prove which
hardware paths the shipped binary uses — UBLKCP (cp.async.bulk), SYNCS (mbarrier), DMMA (FP64 tensor pipe), LDGSTS
(cp.async), DFMA/DMUL/DADD (FP64 pipe), MUFU, IMAD (Philox), LDS/STS, ATOM/RED, and the absence of UTC*MMA / LDTM
(tcgen05 has no f64 kind, so the path cannot use it)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "augmentedgplikelihoods.jl_b200", "libaugcuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s

kern = None
hist = collections.OrderedDict()
archs = collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        archs[m.group(1)] += 1
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1

WATCH = ["UBLKCP", "SYNCS", "DMMA", "LDGSTS", "UTCHMMA", "UTCQMMA", "LDTM", "DFMA", "DMUL", "DADD", "MUFU", "IMAD",
         "LDS", "STS", "LDG", "STG", "ATOM", "ATOMS", "RED", "SHFL", "VOTE", "BAR", "BSSY", "BRA", "WARPSYNC"]
total = collections.Counter()
print(f"# {os.path.relpath(so, ROOT)}: cubins by arch {dict(archs)}; {len(hist)} kernels")
print("# kernel | instructions | " + " ".join(WATCH))
for k, h in hist.items():
    total.update(h)
    n = sum(h.values())
    name = demangle(k)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"\((anonymous namespace\)::)?\w*Args\w*\)$|\([^()]*\)$", "", name).replace("void ", "")
    print(f"{name} | {n} | " + " ".join(f"{w}={h[w]}" for w in WATCH if h[w]))
print("\n# whole library, every opcode")
for op, c in sorted(total.items(), key=lambda t: -t[1]):
    print(f"{op} {c}")
