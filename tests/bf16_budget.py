#!/usr/bin/env python
"""CPU emulation of where bf16 operand rounding costs SNR in the generator (decides which layers
run with split (hi+lo) bf16 activations).  Uses the oracle port; test/analysis tooling only."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import synth  # noqa: E402
from oracle import np_oracle, torch_port  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def rnd(x, mode):
    if mode == "fp32":
        return x
    hi = bf(x)
    if mode == "bf16":
        return hi
    return hi + bf(x - hi)  # "split": two bf16 terms


def forward(w, cfg, mel, mode_of):
    """mode_of(layer_name) -> 'fp32' | 'bf16' | 'split' for the ACTIVATION operand; weights are bf16
    wherever the activation is not fp32."""
    def conv(name, x, **kw):
        m = mode_of(name)
        wt = w[name + ".weight"] if m == "fp32" else bf(w[name + ".weight"])
        return F.conv1d(rnd(x, m), wt, w[name + ".bias"], **kw)

    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        name = f"ups.{i}"
        m = mode_of(name)
        wt = w[name + ".weight"] if m == "fp32" else bf(w[name + ".weight"])
        x = F.conv_transpose1d(rnd(F.leaky_relu(x, 0.1), m), wt, w[name + ".bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            r = x
            for mi, d in enumerate(rd):
                xt = conv(f"{p}.convs1.{mi}", F.leaky_relu(r, 0.1), dilation=d, padding=(rk - 1) * d // 2)
                xt = conv(f"{p}.convs2.{mi}", F.leaky_relu(xt, 0.1), padding=(rk - 1) // 2)
                r = xt + r
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.conv1d(F.leaky_relu(x), w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x).squeeze(1)


def stage_of(name):
    if name.startswith("ups."):
        return int(name.split(".")[1])
    if name.startswith("resblocks."):
        return int(name.split(".")[1]) // 3
    return -1


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    cfg = synth.HIFIGAN_V1
    policies = {
        "all bf16": lambda n: "bf16",
        "ups split": lambda n: "split" if n.startswith("ups") else "bf16",
        "ups split + stage3 split": lambda n: "split" if n.startswith("ups") or stage_of(n) == 3 else "bf16",
        "ups split + stage2,3 split": lambda n: "split" if n.startswith("ups") or stage_of(n) >= 2 else "bf16",
        "ups.2,3 split + stage3 split": lambda n: "split" if n in ("ups.2", "ups.3") or (stage_of(n) == 3 and not n.startswith("ups")) else "bf16",
        "stage3 split only": lambda n: "split" if stage_of(n) == 3 and not n.startswith("ups") else "bf16",
        "all split": lambda n: "split",
    }
    for seed in (1234, 7, 99):
        state = synth.make_state(cfg, seed, "init")
        w = torch_port.fold_state(state)
        mel = torch.from_numpy(synth.make_mel(1, 173, 30))
        with torch.no_grad():
            ref = forward(w, cfg, mel, lambda n: "fp32").numpy()
            for pname, pol in policies.items():
                out = forward(w, cfg, mel, pol).numpy()
                print(f"seed {seed} {pname:32s} SNR {np_oracle.snr_db(ref, out):6.1f} dB  raw {np_oracle.snr_db(ref, out, False):6.1f}  maxabs {np.abs(ref-out).max():.1e}", flush=True)
