#!/usr/bin/env python
"""Bring-up probe for the tcgen05 conv kernel: runs a few shapes through nvse_conv1d_bf16 and
prints the error against an fp32 PyTorch conv on bf16-rounded operands.  Needs a B200."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import conv1d_cl, conv_transpose1d_cl, lib_mod  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def run(cin, cout, k, d, T, B, slope=0.1):
    g = torch.Generator().manual_seed(cin * 1000 + k * 10 + d)
    x = torch.randn((B, cin, T), generator=g).to(dev)
    w = (torch.randn((cout, cin, k), generator=g) / (cin * k) ** 0.5).to(dev)
    b = torch.randn((cout,), generator=g).to(dev)
    ref = F.conv1d(bf(F.leaky_relu(x, slope)), bf(w), b, dilation=d, padding=(k - 1) * d // 2)
    out = conv1d_cl(x, w, b, d, in_slope=slope, tc=True)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    aborted = lib_mod.tc_abort_status()
    print(f"conv Cin={cin} Cout={cout} k={k} d={d} T={T} B={B}: max-abs err {err:.3e} (ref max {ref.abs().max().item():.2f})"
          f"{'  ABORTED' if aborted else ''}", flush=True)
    return err


if __name__ == "__main__":
    print("swap_lbo_sbo =", os.environ.get("NVSE_TC_SWAP_LBO_SBO", "0"), flush=True)
    run(32, 32, 1, 1, 128, 1, slope=1.0)
    run(32, 32, 1, 1, 300, 2)
    run(64, 64, 1, 1, 128, 1)
    run(64, 64, 3, 1, 256, 1)
    run(32, 32, 3, 1, 256, 1)
    run(128, 128, 7, 3, 300, 2)
    run(256, 256, 11, 5, 500, 1)
