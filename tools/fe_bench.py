#!/usr/bin/env python
"""Front-end alone at BASELINE cfg2 (64 x 4 s) and at a >= 1 GB steady-state size: GB/s of algorithmic bytes,
warm (L2-resident input) and cold (L2 flushed between iterations).  ncu-friendly (usage: fe_bench.py [B] [seconds])."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg  # noqa: E402
import synth  # noqa: E402
A = synth.HIFIGAN_V1
dev = "cuda:0"
def mel(y):
    return pkg.mel_spectrogram(y, A["n_fft"], A["num_mels"], A["sampling_rate"], A["hop_size"], A["win_size"], A["fmin"], A["fmax"])
cases = [(int(sys.argv[1]), float(sys.argv[2]))] if len(sys.argv) > 2 else [(64, 4.0), (2048, 4.0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B, sec in cases:
    T = int(sec * 22050)
    y = (torch.rand(B, T, device=dev) * 2 - 1) * 0.5
    out = mel(y); torch.cuda.synchronize()
    nbytes = 4.0 * B * T + 4.0 * out.numel()
    for cold in (False, True):
        ts = []
        for _ in range(10):
            if cold: flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mel(y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        print(f"front-end B={B} T={T} ({nbytes/1e6:.1f} MB algorithmic) {'cold' if cold else 'warm'}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s  "
              f"({100*nbytes/ms/1e6/6545:.1f} % of 6545 GB/s)")
