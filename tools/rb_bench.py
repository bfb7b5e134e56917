#!/usr/bin/env python
"""Time the fused ResBlock1 chain kernel on one shape through the layer-level C ABI; ncu-friendly.
usage: rb_bench.py C k T B [npairs=3] [iters=5]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import lib_mod, stream_ptr  # noqa: E402

Cc, k, T, B = (int(v) for v in sys.argv[1:5])
npairs = int(sys.argv[5]) if len(sys.argv) > 5 else 3
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dils = [int(v) for v in os.environ["RB_DILS"].split(",")] if "RB_DILS" in os.environ else [1, 3, 5][:npairs]
dev = "cuda:0"
lib = lib_mod.load()
x = torch.randn((B, T, Cc), device=dev)
y = torch.zeros((B, T, Cc), device=dev)
ws = [[torch.randn((Cc, Cc, k), device=dev) / (Cc * k) ** 0.5 for _ in range(npairs)] for _ in range(2)]
bs = [[torch.randn((Cc,), device=dev) * 0.1 for _ in range(npairs)] for _ in range(2)]
arr = lambda ts: (C.c_void_p * npairs)(*[t.data_ptr() for t in ts])
a_w1, a_b1, a_w2, a_b2 = arr(ws[0]), arr(bs[0]), arr(ws[1]), arr(bs[1])
dil = (C.c_int * npairs)(*dils)


def run():
    lib_mod.check(lib.nvse_resblock1_bf16(lib_mod.ptr(x), a_w1, a_b1, a_w2, a_b2, dil, npairs, lib_mod.ptr(y), B, T, Cc, k,
                                          1.0 / 3, int(os.environ.get("RB_ACC", "1")), stream_ptr()))


run()
torch.cuda.synchronize()
lib_mod.profile_begin()
for _ in range(iters):
    run()
torch.cuda.synchronize()
for p in lib_mod.profile_end():
    if p["kernel"].startswith(("resblock_tc", "pair_tc")):
        ms = p["ms"] / p["launches"]
        print(f"C={Cc} k={k} T={T} B={B} pairs={npairs}: {ms:.3f} ms  {p['flops'] / p['launches'] / ms / 1e9:.0f} TFLOP/s (algorithmic)  "
              f"{p['bytes'] / p['launches'] / ms / 1e6:.0f} GB/s  aborted={lib_mod.tc_abort_status()}")
