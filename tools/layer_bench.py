#!/usr/bin/env python
"""Time one conv layer shape through the layer-level C ABI (tensor-core path); ncu-friendly.
usage: layer_bench.py C k d T B [residual=1] [accumulate=0] [iters=5]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import lib_mod, stream_ptr  # noqa: E402

C, k, d, T, B = (int(v) for v in sys.argv[1:6])
residual = int(sys.argv[6]) if len(sys.argv) > 6 else 1
accumulate = int(sys.argv[7]) if len(sys.argv) > 7 else 0
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
dev = "cuda:0"
lib = lib_mod.load()
x = torch.randn((B, T, C), device=dev)
w = torch.randn((C, C, k), device=dev) / (C * k) ** 0.5
b = torch.randn((C,), device=dev)
res = torch.randn((B, T, C), device=dev) if residual else None
y = torch.zeros((B, T, C), device=dev)


def run():
    lib_mod.check(lib.nvse_conv1d_bf16(lib_mod.ptr(x), lib_mod.ptr(w), lib_mod.ptr(b), lib_mod.ptr(res), lib_mod.ptr(y),
                                       B, T, C, C, k, d, 0.1, 1.0, accumulate, stream_ptr()))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib_mod.profile_begin()
for _ in range(iters):
    run()
torch.cuda.synchronize()
prof = lib_mod.profile_end()
for p in prof:
    if p["kernel"].startswith("conv_tc"):
        ms = p["ms"] / p["launches"]
        flops = 2.0 * B * T * C * C * k
        byts = B * T * C * 4.0 * (2 + residual + accumulate)
        print(f"C={C} k={k} d={d} T={T} B={B} res={residual} acc={accumulate}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  "
              f"{byts / ms / 1e6:.0f} GB/s (fp32 in/out)  aborted={lib_mod.tc_abort_status()}")
