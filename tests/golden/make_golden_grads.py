#!/usr/bin/env python
"""Generate the backward-pass fixtures tests/golden/grads_*.npz by running the UNMODIFIED
reference generator under torch autograd (what loss_gen_all.backward() does through
HiFiGAN.forward in train_time_wi_inv.py:222-236).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_grads.py

Loss: ``(generator(mel).squeeze(1) * dout).sum()`` with a seeded ``dout`` (tests/synth.make_dout), so every
output sample carries an independent gradient.  Stored: mel, dout, the output, dL/dmel in full, and per
parameter (weight_g / weight_v / bias, in state-dict order) the summary of tests/synth.grad_summary.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (path setup, reference loaders)
import synth  # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    mg._install_librosa_standin()
    only = set(sys.argv[1:])
    ref_hifigan = mg._load("ref_hifigan", "Models/hifigan.py")
    ref_istftnet = mg._load("ref_istftnet", "Models/istftnet.py")
    cases = [
        # name, cfg key, weight seed, regime, B, F, mel seed, dout seed
        ("grads_hifigan_train_f6", "hifigan_train", 51, "unit", 2, 6, 52, 53),
        ("grads_hifigan_train_rb2_f5", "hifigan_train_rb2", 61, "unit", 1, 5, 62, 63),
        ("grads_hifigan_train_init_f4", "hifigan_train", 1234, "init", 1, 4, 72, 73),
        ("grads_istftnet_train_f7", "istftnet_train", 81, "unit", 2, 7, 82, 83),
    ]
    for name, cfg_key, wseed, regime, b, f, mseed, dseed in cases:
        if only and name not in only:
            continue
        cfg = synth.CONFIGS[cfg_key]
        state = synth.make_state(cfg, wseed, regime)
        if cfg["model_name"] == "iSTFTNet":
            gen = mg._ref_generator(ref_istftnet, "iSTFTNet", cfg, state)
        else:
            gen = mg._ref_generator(ref_hifigan, "HiFiGAN", cfg, state)
        gen.train()
        mel = torch.from_numpy(synth.make_mel(b, f, mseed)).requires_grad_(True)
        out = gen(mel)
        out2 = out.squeeze(1) if out.dim() == 3 else out
        dout = synth.make_dout(b, out2.shape[1], dseed)
        (out2 * torch.from_numpy(dout)).sum().backward()
        names, l2, sm, samples = [], [], [], []
        for pname, p in gen.named_parameters():
            a, s, v = synth.grad_summary(p.grad.numpy())
            names.append(pname); l2.append(a); sm.append(s); samples.append(v)
        mg._save(name, {"kind": "grads", "cfg": cfg_key, "weight_seed": wseed, "regime": regime, "mel_seed": mseed,
                        "dout_seed": dseed, "params": names,
                        "ref": f"autograd through the reference {cfg['model_name']}.forward (train_time_wi_inv.py:222-236)"},
                 mel=mel.detach().numpy(), dout=dout, out=out2.detach().numpy(), dmel=mel.grad.numpy(),
                 g_l2=np.array(l2, dtype=np.float64), g_sum=np.array(sm, dtype=np.float64),
                 g_samples=np.stack(samples).astype(np.float32))


if __name__ == "__main__":
    main()
