#!/usr/bin/env python
"""Per-kernel times of one generator + mel-L1 training step (batch 16 x 8192 samples) for either generator.
usage: train_profile.py [hifigan_v1|istftnet] [bf16|fp32]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg, build_generator  # noqa: E402
import synth  # noqa: E402
key = sys.argv[1] if len(sys.argv) > 1 else "istftnet"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
cfg = synth.CONFIGS[key]
a = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda").train()
gen.train_precision = prec
margs = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["sampling_rate"] / 2)
mel_in = torch.from_numpy(synth.make_mel(16, 32, 1)).cuda()
y_mel = pkg.mel_spectrogram(torch.from_numpy(synth.make_wave(16, 32 * 256, 2)).cuda(), *margs)
opt = torch.optim.AdamW(gen.parameters(), 1e-7, betas=(0.8, 0.99))
def step():
    opt.zero_grad(set_to_none=True)
    out = gen(mel_in)
    out = out.squeeze(1) if out.dim() == 3 else out
    (F.l1_loss(y_mel, pkg.mel_spectrogram(out, *margs)) * 45).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print(f"{key} {prec}: {e0.elapsed_time(e1) / 10:.3f} ms per training step (batch 16 x 8192)")
pkg._lib.profile_begin(); step(); torch.cuda.synchronize(); prof = pkg._lib.profile_end()
for k in sorted(prof, key=lambda r: -r["ms"])[:14]:
    print(f"  {k['kernel']:28s} x{k['launches']:3d} {k['ms']:8.3f} ms")
print(f"  sum of kernel times {sum(k['ms'] for k in prof):.3f} ms")
