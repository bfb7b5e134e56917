"""Deterministic synthetic inputs / state dicts shared by the tests, the golden-vector
generator and bench.py.  numpy's PCG64 stream is stable across machines, so the
weights behind a golden fixture are re-created from (cfg, seed) instead of being
committed (HiFi-GAN V1 is 13.9 M parameters)."""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# cfgs/hifigan_v1_config.json:28-45 and cfgs/istftnet_config.json:28-47 (model + audio keys only)
HIFIGAN_V1 = {
    "model_name": "HiFiGAN", "resblock": "1",
    "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4],
    "upsample_initial_channel": 512, "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "segment_size": 16384, "num_mels": 80, "num_freq": 1025, "n_fft": 1024, "hop_size": 256,
    "win_size": 1024, "sampling_rate": 22050, "fmin": 0, "fmax": 8000, "seed": 1234,
}
ISTFTNET = {
    "model_name": "iSTFTNet", "resblock": "1",
    "upsample_rates": [8, 8], "upsample_kernel_sizes": [16, 16],
    "upsample_initial_channel": 512, "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "gen_istft_n_fft": 16, "gen_istft_hop_size": 4,
    "segment_size": 16384, "num_mels": 80, "num_freq": 1025, "n_fft": 1024, "hop_size": 256,
    "win_size": 1024, "sampling_rate": 22050, "fmin": 0, "fmax": 8000, "seed": 1234,
}
# reduced variants: same code paths, small enough for the float64 oracle in seconds
HIFIGAN_SMALL = dict(HIFIGAN_V1, upsample_initial_channel=64)
HIFIGAN_SMALL_RB2 = dict(HIFIGAN_V1, upsample_initial_channel=64, resblock="2",
                         resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])
ISTFTNET_SMALL = dict(ISTFTNET, upsample_initial_channel=64)
# training-path fixtures: every channel count a multiple of 16 (128, 64, 32, 16)
HIFIGAN_TRAIN = dict(HIFIGAN_V1, upsample_initial_channel=256)
HIFIGAN_TRAIN_RB2 = dict(HIFIGAN_SMALL_RB2, upsample_initial_channel=256)
ISTFTNET_TRAIN = dict(ISTFTNET, upsample_initial_channel=256)   # 128, 64 channels

CONFIGS = {
    "hifigan_v1": HIFIGAN_V1, "istftnet": ISTFTNET, "hifigan_small": HIFIGAN_SMALL,
    "hifigan_small_rb2": HIFIGAN_SMALL_RB2, "istftnet_small": ISTFTNET_SMALL,
    "hifigan_train": HIFIGAN_TRAIN, "hifigan_train_rb2": HIFIGAN_TRAIN_RB2,
    "istftnet_train": ISTFTNET_TRAIN,
}


class AttrDict(dict):
    """h as the reference scripts build it (utils.py:11-14)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def layer_specs(cfg):
    """(state-dict prefix, kind, Cin, Cout, k) for every conv of the generator, in the
    order the reference constructors create them (hifigan.py:84-106, istftnet.py:272-297)."""
    c0 = cfg["upsample_initial_channel"]
    specs = [("conv_pre", "conv", 80, c0, 7)]
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        specs.append((f"ups.{i}", "convT", c0 // 2 ** i, c0 // 2 ** (i + 1), k))
    n_k = len(cfg["resblock_kernel_sizes"])
    for i in range(len(cfg["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            p = f"resblocks.{i * n_k + j}"
            if str(cfg["resblock"]) == "1":
                specs += [(f"{p}.convs1.{m}", "conv", ch, ch, rk) for m in range(len(rd))]
                specs += [(f"{p}.convs2.{m}", "conv", ch, ch, rk) for m in range(len(rd))]
            else:
                specs += [(f"{p}.convs.{m}", "conv", ch, ch, rk) for m in range(len(rd))]
    out_ch = 1 if cfg["model_name"] == "HiFiGAN" else cfg["gen_istft_n_fft"] + 2
    specs.append(("conv_post", "conv", ch, out_ch, 7))
    return specs


def make_state(cfg, seed, regime="init", weight_norm=True):
    """Synthetic generator state dict in the reference's checkpoint key format
    (``*.weight_g`` / ``*.weight_v`` / ``*.bias``; train_time_wi_inv.py:254).

    regime "init": weight_v ~ N(0, 0.01) like init_weights (hifigan.py:9-12), so the
                   output is bias/DC dominated as at the reference's random init.
    regime "unit": variance-preserving weights so every layer matters numerically.
    weight_g is perturbed away from ||v|| so the fold is exercised."""
    rng = np.random.default_rng(seed)
    state = {}
    for prefix, kind, cin, cout, k in layer_specs(cfg):
        shape = (cout, cin, k) if kind == "conv" else (cin, cout, k)
        fan_in = cin * k if kind == "conv" else cin * k / max(1, k // 2)
        if regime == "init" and prefix != "conv_pre":
            std = 0.01
        else:
            std = 0.7 / np.sqrt(fan_in)
            if regime == "unit" and ".convs" in prefix:
                std *= 0.6
        v = rng.normal(0.0, std, size=shape).astype(np.float32)
        bound = 1.0 / np.sqrt(cin * k)
        b = rng.uniform(-bound, bound, size=(cout,)).astype(np.float32)
        if weight_norm:
            norm = np.sqrt((v.astype(np.float64) ** 2).reshape(shape[0], -1).sum(1))
            g = (norm * rng.uniform(0.8, 1.2, size=shape[0])).astype(np.float32).reshape(shape[0], 1, 1)
            state[prefix + ".weight_g"] = g
            state[prefix + ".weight_v"] = v
        else:
            state[prefix + ".weight"] = v
        state[prefix + ".bias"] = b
    return state


def make_wave(batch, samples, seed):
    """uniform +-0.5 float32 white noise."""
    rng = np.random.default_rng(seed)
    return ((rng.random((batch, samples), dtype=np.float32) * 2 - 1) * 0.5).astype(np.float32)


def make_mel(batch, frames, seed):
    """log-mel-like input in the range real log-mels occupy ([-11.5, 2])."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-6.0, 1.0, size=(batch, 80, frames)).astype(np.float32)


def make_dout(batch, samples, seed):
    """dL/d(waveform) of a synthetic loss ``(wav * dout).sum()``."""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, 1.0, size=(batch, samples)).astype(np.float32)


GRAD_SAMPLES = 8


def grad_summary(g):
    """(l2 norm, sum, GRAD_SAMPLES evenly spaced entries) of one gradient tensor -- what the backward
    fixtures store per parameter (full gradients of 3.5 M parameters would be 14 MB)."""
    f = np.asarray(g, dtype=np.float64).reshape(-1)
    idx = np.linspace(0, f.size - 1, GRAD_SAMPLES).astype(np.int64)
    return float(np.sqrt((f * f).sum())), float(f.sum()), f[idx].astype(np.float32)


def load_golden(name):
    path = os.path.join(HERE, "golden", name + ".npz")
    with np.load(path, allow_pickle=False) as z:
        out = {k: z[k] for k in z.files}
    if "meta" in out:
        out["meta"] = json.loads(str(out["meta"]))
    return out


def mel_mismatch(out, ref, rtol=2e-5, atol=1e-6):
    """Worst-case ratio of the linear-domain error of two log-mels to its tolerance
    ``rtol*exp(ref) + atol*max(exp(ref))``; <= 1 passes.  Compared in the linear domain
    because log() near the 1e-5 clamp floor (dataset.py:27-28) amplifies fp32 rounding
    of the mel sum by up to 1e5 (the reference itself is only defined to fp32 there)."""
    lo = np.exp(np.asarray(out, dtype=np.float64))
    lr = np.exp(np.asarray(ref, dtype=np.float64))
    return float((np.abs(lo - lr) / (rtol * lr + atol * lr.max())).max())
