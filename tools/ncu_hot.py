#!/usr/bin/env python
"""Hottest SASS instructions of a kernel in an .ncu-rep (warp stall samples), with the stall reasons that dominate each.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ai, ii = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(k, c) for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[2:] if len(r) == len(hdr)]
total = sum(int(r[ai]) for r in body)
tot_inst = sum(int(r[ii]) for r in body)
print(f"# {rows[0][1]}\n# {len(body)} SASS instructions, {total} stall samples, {tot_inst} warp-instructions executed")
agg = {}
for r in body:
    for k, c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[k] or 0)
print("# stall reasons overall: " + ", ".join(f"{c[6:]} {v / max(total, 1):.2f}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][ai]))[:top]:
    reasons = sorted(((int(r[k] or 0), c[6:]) for k, c in stall_cols), reverse=True)[:3]
    print(f"{int(r[ai]) / max(total, 1):6.3f}  #{n:5d} exec={int(r[ii]):9d}  {r[si].strip()[:70]:70s}  " + " ".join(f"{c}:{v}" for v, c in reasons if v))
