#!/usr/bin/env python
"""Per-kernel times of one HiFi-GAN V1 forward at the cfg5 micro-batch (32 x 862 frames), grouped by kernel name.
usage: ups_bench.py [pattern]   (pattern: substring of the kernel names to print, default conv_tc)"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import lib_mod, pkg  # noqa: E402
pat = sys.argv[1] if len(sys.argv) > 1 else "_tc["
import synth  # noqa: E402
torch.manual_seed(0)
g = pkg.HiFiGAN(synth.AttrDict(synth.HIFIGAN_V1)).to("cuda").eval()
g.remove_weight_norm()
mel = torch.randn(32, 80, 862, device="cuda")
with torch.no_grad():
    for _ in range(2): g(mel)
    torch.cuda.synchronize()
    lib_mod.profile_begin()
    for _ in range(3): g(mel)
    torch.cuda.synchronize()
tot = 0.0
for p in lib_mod.profile_end():
    ms = p["ms"] / 3
    tot += ms
    if pat in p["kernel"] or pat == "all":
        print(f"{p['kernel']:28s} {ms:7.3f} ms/forward  {p['flops'] / 3 / ms / 1e9:7.0f} TF/s  {p['bytes'] / 3 / ms / 1e6:7.0f} GB/s")
print(f"total {tot:.3f} ms/forward")
