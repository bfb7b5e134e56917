#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):    python tests/golden/make_golden.py

The reference modules are loaded by file path (never ``import Models`` -- its
``__init__`` pulls matplotlib/librosa).  ``librosa`` is not
installed here, so a ``sys.modules`` stand-in supplies ``librosa.filters.mel`` from
``oracle.np_oracle.mel_filterbank`` (published-algorithm restatement; "parity
unpinned" for that one function) plus the three ``librosa.util`` names that
``Models/istftnet.py`` imports at module level but does not use on the hot path.

Weights are NOT stored: fixtures record (config name, seed, regime) and tests
rebuild the state dict with ``tests/synth.make_state``.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from oracle import np_oracle  # noqa: E402

REF = os.environ.get("NVSE_REFERENCE", "/root/reference")


def _install_librosa_standin():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")

    def pad_center(data, *, size, axis=-1, **kw):
        n = data.shape[axis]
        lp = (size - n) // 2
        widths = [(0, 0)] * data.ndim
        widths[axis] = (lp, size - n - lp)
        return np.pad(data, widths)

    util.pad_center = pad_center
    util.tiny = lambda x: np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny
    util.normalize = lambda x, **kw: x / max(1e-12, np.abs(x).max())
    filters.mel = lambda *, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw: np_oracle.mel_filterbank(
        sr, n_fft, n_mels, fmin, fmax)
    lib.util, lib.filters = util, filters
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filters})


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _save(name, meta, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"wrote {path}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items()))


def _ref_generator(mod, cls_name, cfg, state):
    h = synth.AttrDict(cfg)
    torch.manual_seed(cfg["seed"])
    gen = getattr(mod, cls_name)(h)
    missing = gen.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    gen.eval()
    return gen


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    _install_librosa_standin()
    ref_dataset = _load("ref_dataset", "dataset.py")
    ref_hifigan = _load("ref_hifigan", "Models/hifigan.py")
    ref_istftnet = _load("ref_istftnet", "Models/istftnet.py")

    # ---- mel_spectrogram (dataset.py:53-91) ------------------------------------
    a = synth.HIFIGAN_V1
    cases = {
        "mel_b2_t4100": (synth.make_wave(2, 4100, 11), a["fmax"]),       # T not a multiple of hop
        "mel_b1_t513": (synth.make_wave(1, 513, 12), a["fmax"]),         # smallest legal T (> n_fft/2)
        "mel_b3_t8192_fmax_half": (synth.make_wave(3, 8192, 13), a["sampling_rate"] / 2),  # train_time_wi_inv.py:185
        "mel_1d_t22050": (synth.make_wave(1, 22050, 14)[0], a["fmax"]),  # 1-D input
    }
    # a tonal + silent signal so the 1e-5 clamp floor is hit (dataset.py:27-28)
    t = np.arange(6000, dtype=np.float64)
    tone = (0.3 * np.sin(2 * np.pi * 440.0 * t / 22050)).astype(np.float32)
    tone[3000:] = 0.0
    cases["mel_tone_silence"] = (tone[None], a["fmax"])
    for name, (y, fmax) in cases.items():
        ref_dataset.mel_window.clear()
        out = ref_dataset.mel_spectrogram(torch.from_numpy(y), a["n_fft"], a["num_mels"], a["sampling_rate"],
                                          a["hop_size"], a["win_size"], a["fmin"], fmax)
        _save(name, {"kind": "mel", "fmax": fmax, "ref": "dataset.py:53-91"}, y=y, out=out.numpy())

    # ---- generators ---------------------------------------------------------------
    gen_cases = [
        # name, cfg key, module, class, weight seed, regime, B, F, mel seed
        ("hifigan_v1_init_f6", "hifigan_v1", ref_hifigan, "HiFiGAN", 1234, "init", 1, 6, 21),
        ("hifigan_v1_unit_f9", "hifigan_v1", ref_hifigan, "HiFiGAN", 7, "unit", 2, 9, 22),
        ("hifigan_small_unit_f33", "hifigan_small", ref_hifigan, "HiFiGAN", 99, "unit", 2, 33, 23),
        ("hifigan_small_f1", "hifigan_small", ref_hifigan, "HiFiGAN", 5, "unit", 1, 1, 24),
        ("hifigan_small_rb2_f17", "hifigan_small_rb2", ref_hifigan, "HiFiGAN", 3, "unit", 1, 17, 25),
        ("istftnet_init_f5", "istftnet", ref_istftnet, "iSTFTNet", 1234, "init", 1, 5, 26),
        ("istftnet_unit_f7", "istftnet", ref_istftnet, "iSTFTNet", 8, "unit", 2, 7, 27),
        ("istftnet_small_unit_f40", "istftnet_small", ref_istftnet, "iSTFTNet", 9, "unit", 2, 40, 28),
    ]
    for name, cfg_key, mod, cls, wseed, regime, b, f, mseed in gen_cases:
        cfg = synth.CONFIGS[cfg_key]
        state = synth.make_state(cfg, wseed, regime)
        gen = _ref_generator(mod, cls, cfg, state)
        mel = synth.make_mel(b, f, mseed)
        with torch.no_grad():
            out_wn = gen(torch.from_numpy(mel)).numpy()
            gen.remove_weight_norm()
            out = gen(torch.from_numpy(mel)).numpy()
        assert np.abs(out - out_wn).max() < 1e-5
        _save(name, {"kind": "generator", "cfg": cfg_key, "weight_seed": wseed, "regime": regime,
                     "mel_seed": mseed, "ref": f"Models/{mod.__name__[4:]}.py {cls}.forward"},
              mel=mel, out=out)

    # ---- constructor parity: state-dict keys, shapes and seeded-init checksums ----------
    for cfg_key, mod, cls in (("hifigan_v1", ref_hifigan, "HiFiGAN"), ("istftnet", ref_istftnet, "iSTFTNet"),
                              ("hifigan_small_rb2", ref_hifigan, "HiFiGAN")):
        cfg = synth.CONFIGS[cfg_key]
        torch.manual_seed(cfg["seed"])
        gen = getattr(mod, cls)(synth.AttrDict(cfg))
        sd = gen.state_dict()
        table = {k: {"shape": list(v.shape), "sum": float(v.double().sum()), "abs": float(v.double().abs().sum())}
                 for k, v in sd.items()}
        with open(os.path.join(HERE, f"state_{cfg_key}.json"), "w") as f:
            json.dump({"ref": f"{cls}.__init__ under torch.manual_seed({cfg['seed']})", "tensors": table}, f, indent=0)
        print(f"wrote state_{cfg_key}.json ({len(table)} tensors)")

    # ---- iSTFT head alone (istftnet.py:183-188) ----------------------------------
    rng = np.random.default_rng(31)
    mag = np.exp(rng.normal(0, 1, size=(2, 9, 37))).astype(np.float32)
    pha = np.sin(rng.normal(0, 2, size=(2, 9, 37))).astype(np.float32)
    head = ref_istftnet.TorchSTFT(filter_length=16, hop_length=4, win_length=16)
    out = head.inverse(torch.from_numpy(mag), torch.from_numpy(pha)).squeeze(1).numpy()
    _save("istft_head_t37", {"kind": "istft", "ref": "Models/istftnet.py:183-188"}, mag=mag, phase=pha, out=out)

    # ---- end to end wav -> mel -> wav, as infers/inference_hifigan.py:82-84 --------
    cfg = synth.HIFIGAN_SMALL
    state = synth.make_state(cfg, 41, "unit")
    gen = _ref_generator(ref_hifigan, "HiFiGAN", cfg, state)
    y = synth.make_wave(1, 5000, 42)
    ref_dataset.mel_window.clear()
    with torch.no_grad():
        mel = ref_dataset.mel_spectrogram(torch.from_numpy(y), cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"],
                                          cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["fmax"])
        out = gen(mel).numpy()
    _save("e2e_hifigan_small_t5000", {"kind": "e2e", "cfg": "hifigan_small", "weight_seed": 41, "regime": "unit",
                                      "ref": "infers/inference_hifigan.py:82-84"}, y=y, mel=mel.numpy(), out=out)


if __name__ == "__main__":
    main()
