"""Shared test helpers: package loader and ctypes wrappers over the layer-level C ABI."""
from __future__ import annotations

import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib_mod = pkg._lib


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv1d_cl(x_bct, w, b, dilation, in_slope=1.0, residual_bct=None, out_scale=1.0, y0_bct=None, tc=False):
    """Run nvse_conv1d_{f32,bf16} on a [B, C, T] tensor (PyTorch layout) and return [B, Cout, T]."""
    import torch
    lib = lib_mod.load()
    x = x_bct.transpose(1, 2).contiguous()
    B, T, Cin = x.shape
    Cout, _, k = w.shape
    res = residual_bct.transpose(1, 2).contiguous() if residual_bct is not None else None
    if y0_bct is not None:
        y = y0_bct.transpose(1, 2).contiguous().clone()  # never alias the caller's y0
        acc = 1
    else:
        y = torch.empty((B, T, Cout), dtype=torch.float32, device=x.device)
        acc = 0
    fn = lib.nvse_conv1d_bf16 if tc else lib.nvse_conv1d_f32
    lib_mod.check(fn(lib_mod.ptr(x), lib_mod.ptr(w.contiguous()), lib_mod.ptr(b), lib_mod.ptr(res), lib_mod.ptr(y),
                     B, T, Cin, Cout, k, dilation, in_slope, out_scale, acc, stream_ptr()))
    return y.transpose(1, 2).contiguous()


def conv_transpose1d_cl(x_bct, w, b, stride, padding, in_slope=1.0, tc=False):
    import torch
    lib = lib_mod.load()
    x = x_bct.transpose(1, 2).contiguous()
    B, T, Cin = x.shape
    _, Cout, k = w.shape
    Tout = (T - 1) * stride - 2 * padding + k
    y = torch.empty((B, Tout, Cout), dtype=torch.float32, device=x.device)
    fn = lib.nvse_conv_transpose1d_bf16 if tc else lib.nvse_conv_transpose1d_f32
    lib_mod.check(fn(lib_mod.ptr(x), lib_mod.ptr(w.contiguous()), lib_mod.ptr(b), lib_mod.ptr(y), B, T, Cin, Cout, k,
                     stride, padding, in_slope, stream_ptr()))
    return y.transpose(1, 2).contiguous()


def resblock1_cl(x_bct, w1, b1, w2, b2, dilations, out_scale=1.0, y0_bct=None):
    """Run nvse_resblock1_bf16 (fused ResBlock1 chain) on a [B, C, T] tensor; returns [B, C, T]."""
    import torch
    lib = lib_mod.load()
    x = x_bct.transpose(1, 2).contiguous()
    B, T, Cc = x.shape
    k = w1[0].shape[2]
    n = len(dilations)
    if y0_bct is not None:
        y, acc = y0_bct.transpose(1, 2).contiguous().clone(), 1
    else:
        y, acc = torch.full((B, T, Cc), float("nan"), dtype=torch.float32, device=x.device), 0
    keep = [[t.contiguous() for t in ts] for ts in (w1, b1, w2, b2)]
    arrs = [(C.c_void_p * n)(*[t.data_ptr() for t in ts]) for ts in keep]
    dil = (C.c_int * n)(*dilations)
    lib_mod.check(lib.nvse_resblock1_bf16(lib_mod.ptr(x), arrs[0], arrs[1], arrs[2], arrs[3], dil, n, lib_mod.ptr(y),
                                          B, T, Cc, k, out_scale, acc, stream_ptr()))
    torch.cuda.synchronize()
    return y.transpose(1, 2).contiguous()


def report(line):
    """Print a measurement and, on the GPU box, also append it to gpurun_out/parity_report.txt."""
    print(line)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_report.txt"), "a") as f:
            f.write(line + "\n")


def build_generator(cfg, state, device, remove_wn=False):
    """The product module for ``cfg`` loaded from a reference-format state dict."""
    import torch
    import synth
    cls = pkg.HiFiGAN if cfg["model_name"] == "HiFiGAN" else pkg.iSTFTNet
    gen = cls(synth.AttrDict(cfg))
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=True)
    gen = gen.to(device).eval()
    if remove_wn:
        gen.remove_weight_norm()
    return gen
