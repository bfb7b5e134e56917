#!/usr/bin/env python
"""One training step (after warm-up) of the generator + mel-L1 at the reference's shape: the ncu target for the
backward kernels.  usage: train_once.py [precision=bf16] [steps=1]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator, pkg  # noqa: E402
import synth  # noqa: E402
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda").train()
gen.precision = prec
margs = (cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["sampling_rate"] / 2)
mel_in = torch.from_numpy(synth.make_mel(16, 32, 1)).cuda()
y_mel = pkg.mel_spectrogram(torch.from_numpy(synth.make_wave(16, 8192, 2)).cuda(), *margs)
for _ in range(steps + 1):
    gen.zero_grad(set_to_none=True)
    loss = F.l1_loss(y_mel, pkg.mel_spectrogram(gen(mel_in), *margs)) * 45
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()), "tc_abort", pkg._lib.tc_abort_status())
