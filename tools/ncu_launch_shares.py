#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`).  usage: ncu_launch_shares.py x.csv [title] [comma-separated kernel-name substrings to leave out]"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
skip = [x for x in (sys.argv[3].split(',') if len(sys.argv) > 3 else []) if x]
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
    name = r[ki].replace("nvse::<unnamed>::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    name = name.split("(")[0]
    if any(x in name for x in skip):
        continue
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
if len(sys.argv) > 2: print(sys.argv[2])
print(f"total {tot:.2f} ms over {n} launches\n")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * v[1] / tot:6.2f}%  x{v[0]:4d} {v[1]:10.3f} ms  {k[:90]}")
tc = sum(v[1] for k, v in agg.items() if any(s in k for s in ("resblock_", "pair_tc", "conv_tc", "ups_tc")))
print(f"\ntensor-core kernels (resblock_* + pair_tc + ups_tc + conv_tc): {100 * tc / tot:.1f} % of device time under ncu")
