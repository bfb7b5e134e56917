#!/usr/bin/env python
"""Per-kernel times of one iSTFTNet forward at the cfg4 shape (32 x 690 frames), 16-bit tensor-core path."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg, build_generator  # noqa: E402
import synth  # noqa: E402
cfg = synth.ISTFTNET
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda", remove_wn=True)
gen.precision = "bf16"
B, F = int(os.environ.get("B", 32)), int(os.environ.get("F", 690))
mel = torch.from_numpy(synth.make_mel(B, F, 1)).cuda()
with torch.no_grad():
    for _ in range(3): gen(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gen(mel)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    pkg._lib.profile_begin()
    gen(mel); torch.cuda.synchronize()
    prof = pkg._lib.profile_end()
flop = 4.116e8 * B * F
print(f"iSTFTNet {B} x {F} frames: {ms:.3f} ms per forward = {B * F * 256 / 22050 / ms * 1e3:.0f} audio-s/s, {flop / ms / 1e9:.0f} TFLOP/s algorithmic")
for k in sorted(prof, key=lambda r: -r["ms"]):
    print(f"  {k['kernel']:28s} x{k['launches']:3d} {k['ms']:8.3f} ms  {k['flops'] / max(k['ms'], 1e-9) / 1e9:8.1f} TF/s {k['bytes'] / max(k['ms'], 1e-9) / 1e6:8.0f} GB/s")
