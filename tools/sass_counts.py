#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (runs where nvcc's cuobjdump is, no GPU needed):
usage: tools/sass_counts.py [libnvse_b200.so] > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                         "neural-vocoders-as-speech-enhancers_b200", "csrc", "libnvse_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", "UTCHMMA.WS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "UTCBAR", "HMMA", "LDGSTS", "FFMA"]
counts, order, cur, i = {}, [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[i].replace("nvse::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        while cur in counts:
            cur += "'"
        i += 1
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?\S+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur][base] += 1
        if op.startswith("UTCHMMA") and ".WS" in op:
            counts[cur]["UTCHMMA.WS"] += 1
print("SASS mnemonic counts per kernel of libnvse_b200.so (cuobjdump -sass; sm_100a).  UTCHMMA = tcgen05.mma (.WS = weight-stationary form),\n"
      "LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, LDGSTS = cp.async;\n"
      "no legacy HMMA anywhere.  FFMA: the fp32 CUDA-core kernels (front-end, fp32 parity path, discriminators).\n")
print(f"{'kernel':80s}" + "".join(f"{c:>11s}" for c in COLS))
tot = collections.Counter()
for k in order:
    print(f"{k[:80]:80s}" + "".join(f"{counts[k][c]:11d}" for c in COLS))
    tot.update({c: counts[k][c] for c in COLS})
print(f"{'TOTAL':80s}" + "".join(f"{tot[c]:11d}" for c in COLS))
