#!/usr/bin/env python
"""profiles/r01_ncu_traffic.json from an `ncu --set full` report of one forward of bench.py:
mean DRAM bytes per launch over the tensor-core kernels, next to their algorithmic bytes.
usage: ncu_traffic.py report.ncu-rep bench.json micro_batch seconds"""
import csv, io, json, subprocess, sys
rep, bench_json, mb, sec = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
per_kernel, tot, n = {}, 0.0, 0
for r in data:
    name = r[ix["Kernel Name"]]
    if not any(s in name for s in ("resblock_", "pair_tc", "conv_tc", "ups_tc")):
        continue
    # ncu scales units per row in the raw page: re-read them from the per-row unit columns when present
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    key = name.split("<")[0].split("::")[-1] + "<" + name.split("<")[1].split(">")[0] + ">"
    a = per_kernel.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += rd + wr
    tot += rd + wr; n += 1
bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
steps = bench["steps"]
tc = [k for k in bench["kernels"] if k["kernel"].startswith(("conv_tc", "ups_tc", "resblock_tc", "pair_tc"))]
alg = sum(k["gbs"] * 1e9 * k["ms_per_step"] * 1e-3 for k in tc)  # algorithmic bytes per step (per forward x forwards per step)
launches_step = sum(k["launches"] for k in tc) / steps
out = {"micro_batch": mb, "seconds": sec, "launches": n, "dram_bytes_per_launch": tot / n,
       "algorithmic_bytes_per_launch": alg / launches_step, "per_kernel": {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0]} for k, v in per_kernel.items()}}
print(json.dumps(out, indent=1))
