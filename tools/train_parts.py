#!/usr/bin/env python
"""Where a training step's device time goes: CUDA-event timing of its parts (tensor-core path)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator, pkg  # noqa: E402
import synth  # noqa: E402
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda").train()
gen.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
margs = (cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["sampling_rate"] / 2)
mel_in = torch.from_numpy(synth.make_mel(16, 32, 1)).cuda()
y_mel = pkg.mel_spectrogram(torch.from_numpy(synth.make_wave(16, 8192, 2)).cuda(), *margs)
opt = torch.optim.AdamW(gen.parameters(), 1e-7, betas=(0.8, 0.99))
N = 10
acc = {}
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(N + 3):
    t = [ev()]
    opt.zero_grad(set_to_none=True); t.append(ev())
    y_g = gen(mel_in); t.append(ev())
    loss = F.l1_loss(y_mel, pkg.mel_spectrogram(y_g, *margs)) * 45; t.append(ev())
    loss.backward(); t.append(ev())
    opt.step(); t.append(ev())
    torch.cuda.synchronize()
    if it >= 3:
        for name, a, b in zip(("zero_grad", "load+forward_train", "mel+loss", "backward (mel, generator, weight_norm)", "AdamW"), t[:-1], t[1:]):
            acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) / N
print({k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3))
