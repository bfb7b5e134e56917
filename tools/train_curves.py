#!/usr/bin/env python
"""Loss curves of the generator + 45 * mel-L1 objective (train_time_wi_inv.py:73,166-179,231-236) at the reference's training
shape (HiFi-GAN V1, batch 16 x 8192 samples, AdamW 2e-4 / (0.8, 0.99)) over a small fixed corpus of synthetic segments:
the fp32 training path against the tensor-core path (bf16 operands) from the same initial weights and the same batches.
usage: train_curves.py [steps=300]"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg, build_generator, lib_mod  # noqa: E402
import synth  # noqa: E402
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg = synth.HIFIGAN_V1
a = cfg
margs = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["sampling_rate"] / 2)
margs_in = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["fmax"])
# 8 batches of 16 segments: band-limited noise with a few tones, so the mel targets have structure to fit
rng = np.random.default_rng(0)
batches = []
for b in range(8):
    t = np.arange(8192) / 22050.0
    y = 0.2 * synth.make_wave(16, 8192, 500 + b)
    for r in range(16):
        for f0 in rng.uniform(100, 3000, size=3):
            y[r] += 0.15 * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28))
    y = torch.from_numpy(y.astype(np.float32)).cuda()
    batches.append((pkg.mel_spectrogram(y, *margs_in)[:, :, :32].contiguous(), pkg.mel_spectrogram(y, *margs)))
curves = {}
for prec in ("fp32", "bf16"):
    gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda").train()
    gen.train_precision = prec
    opt = torch.optim.AdamW(gen.parameters(), 2e-4, betas=(0.8, 0.99))
    losses = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        x, y_mel = batches[s % len(batches)]
        opt.zero_grad(set_to_none=True)
        y_g = gen(x)
        n = min(y_g.shape[-1], 8192)
        loss = F.l1_loss(y_mel[:, :, :1 + n // 256], pkg.mel_spectrogram(y_g[..., :n], *margs)) * 45
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    e1.record(); torch.cuda.synchronize()
    curves[prec] = torch.stack(losses).cpu().numpy()
    print(f"{prec}: {e0.elapsed_time(e1) / steps:.2f} ms per step; loss {curves[prec][0]:.3f} -> {curves[prec][-8:].mean():.3f} (mean of the last 8 steps)")
assert not lib_mod.tc_abort_status()
f, b = curves["fp32"], curves["bf16"]
k = 8
fs = np.array([f[i:i + k].mean() for i in range(0, steps - k + 1, k)])   # one value per pass over the corpus
bs = np.array([b[i:i + k].mean() for i in range(0, steps - k + 1, k)])
print("step   fp32-loss  tensor-core-loss  relative difference   (means over one pass of the 8 batches)")
for i in range(0, len(fs), max(1, len(fs) // 12)):
    print(f"{i * k:5d}  {fs[i]:9.4f}  {bs[i]:9.4f}        {(bs[i] - fs[i]) / fs[i]:+.3%}")
print(f"worst per-pass relative difference {np.abs(bs - fs).max() / 1:.4f} absolute, {np.abs((bs - fs) / fs).max():.3%} relative; final pass {(bs[-1] - fs[-1]) / fs[-1]:+.3%}")
