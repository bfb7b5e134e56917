#!/usr/bin/env python
"""Top SASS instructions of one kernel of an .ncu-rep by warp-stall samples (source page).  usage: ncu_hot_sass.py rep kernel_index [n]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
h = rows[1]
si, ni = h.index("Source"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) > ni and r[ni].isdigit()]
tot = sum(int(r[ni] or 0) for r in data)
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ni] or 0))
for i in order[:n]:
    r = data[i]
    top = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{100.0 * int(r[ni]) / tot:5.1f}%  #{i:5d}  {r[si].strip()[:90]:90s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
