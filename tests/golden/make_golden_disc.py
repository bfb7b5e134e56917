#!/usr/bin/env python
"""Generate the discriminator fixtures tests/golden/disc_*.npz and state_disc.json by running the UNMODIFIED reference
classes (Models/models.py: MultiPeriodDiscriminator :90-113, MultiScaleDiscriminator :217-246) and the reference losses
the trainer applies to them (train_time_wi_inv.py:188-236) under torch autograd, on CPU.

Run in the build container only (needs /root/reference):    python tests/golden/make_golden_disc.py

The weights are the reference constructors' own random initialisation under ``torch.manual_seed(seed)``; the drop-in
replays the same draws, which state_disc.json pins (per-tensor checksums), so the fixtures need not store 70 M
parameters.  Stored per case: the two input batches, every logit tensor in full, a summary of every feature map and of
every parameter gradient, and dL/dy_hat in full, for
    L_D = ls_discriminator_loss(D(y), D(y_hat))                 (discriminator step)
    L_G = ls_generator_loss(D(y_hat)) + feature_loss(fmaps)     (generator step)."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import synth  # noqa: E402

MPD_RESHAPES = [2, 3, 5, 7, 11]  # cfgs/hifigan_v1_config.json: mpd_reshapes


def load_reference_models():
    mg._install_librosa_standin()
    sys.path.insert(0, os.path.join(mg.ROOT, "tests", "shims"))  # matplotlib stand-in for utils.py:3-7
    sys.modules.setdefault("utils", mg._load("utils", "utils.py"))
    sys.modules.setdefault("dataset", mg._load("dataset", "dataset.py"))
    return mg._load("ref_models", "Models/models.py")


def run_case(ref, kind, seed, y, y_hat):
    torch.manual_seed(seed)
    net = ref.MultiPeriodDiscriminator(MPD_RESHAPES) if kind == "mpd" else ref.MultiScaleDiscriminator()
    net.train()
    state_sums = {k: synth.grad_summary(v.detach().numpy())[:2] for k, v in net.state_dict().items()}
    y_t = torch.from_numpy(y)
    yh = torch.from_numpy(y_hat).requires_grad_(True)
    # discriminator step
    d_r, d_g, _, _ = net(y_t, yh.detach())
    loss_d, _, _ = ref.ls_discriminator_loss(d_r, d_g)
    loss_d.backward()
    gd = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    net.zero_grad()
    # generator step (second forward: spectral_norm's power iteration advances again, as in the trainer)
    d_r2, d_g2, f_r, f_g = net(y_t, yh)
    loss_g = ref.ls_generator_loss(d_g2)[0] + ref.feature_loss(f_r, f_g)
    loss_g.backward()
    names = [n for n, _ in net.named_parameters()]
    arrays = {"y": y, "y_hat": y_hat, "loss_d": np.array(loss_d.item()), "loss_g": np.array(loss_g.item()), "dyhat": yh.grad.numpy()}
    for i, (a, b) in enumerate(zip(d_r, d_g)):
        arrays[f"dr{i}"], arrays[f"dg{i}"] = a.detach().numpy(), b.detach().numpy()
    for i, (a, b) in enumerate(zip(d_r2, d_g2)):
        arrays[f"dr2_{i}"], arrays[f"dg2_{i}"] = a.detach().numpy(), b.detach().numpy()
    fm = [synth.grad_summary(t.detach().numpy()) for fr in f_g for t in fr]
    arrays["fmap_l2"] = np.array([a for a, _, _ in fm]); arrays["fmap_sum"] = np.array([s for _, s, _ in fm])
    arrays["fmap_samples"] = np.stack([v for _, _, v in fm])
    for tag, grads in (("gd", gd), ("gg", {n: p.grad for n, p in net.named_parameters()})):
        sm = [synth.grad_summary(grads[n].numpy()) for n in names]
        arrays[tag + "_l2"] = np.array([a for a, _, _ in sm]); arrays[tag + "_sum"] = np.array([s for _, s, _ in sm])
        arrays[tag + "_samples"] = np.stack([v for _, _, v in sm])
    return names, state_sums, arrays


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference_models()
    cases = [
        # name, kind, ctor seed, B, T, wave seeds
        ("disc_mpd_b2_t2200", "mpd", 1234, 2, 2200, 31, 32),   # 2200 = 2^3 * 5^2 * 11: periods 3 and 7 need the reflect pad
        ("disc_msd_b2_t2048", "msd", 1234, 2, 2048, 33, 34),
        ("disc_msd_b1_t999", "msd", 77, 1, 999, 35, 36),         # odd lengths through the strided layers and the mean pools
    ]
    states = {}
    for name, kind, seed, b, t, s1, s2 in cases:
        y, y_hat = synth.make_wave(b, t, s1), synth.make_wave(b, t, s2)
        names, sums, arrays = run_case(ref, kind, seed, y, y_hat)
        states[f"{kind}_seed{seed}"] = {k: [float(a), float(s)] for k, (a, s) in sums.items()}
        mg._save(name, {"kind": kind, "seed": seed, "params": names, "mpd_reshapes": MPD_RESHAPES,
                        "ref": "Models/models.py:15-113,187-246 + train_time_wi_inv.py:188-236 under autograd"}, **arrays)
    with open(os.path.join(HERE, "state_disc.json"), "w") as f:
        json.dump(states, f, indent=0, sort_keys=True)
    print("wrote state_disc.json:", {k: len(v) for k, v in states.items()})


if __name__ == "__main__":
    main()
