"""Host-side logic and ABI surface; runs without a GPU."""
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import synth
from oracle import np_oracle
from util import ROOT, lib_mod, pkg

HEADER = os.path.join(ROOT, "include", "nvse_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NVSE_API[^;(]*?\b(nvse_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(lib_mod.PROTOTYPES) == declared          # the binding covers the header exactly
    lib = lib_mod.load()                                     # dlopen + getattr of each symbol
    assert lib.nvse_abi_version() == lib_mod.ABI_VERSION
    nm = subprocess.run(["nm", "-D", "--defined-only", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (nvse_\w+)", nm)))
    assert exported == declared                              # nothing undeclared leaks out either


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("fmax", [8000, 8000.0, 11025.0, None])
def test_mel_basis_matches_oracle_bit_for_bit(fmax):
    mine = pkg.melbasis.slaney_mel_basis(22050, 1024, 80, 0, fmax)
    ref = np_oracle.mel_filterbank(22050, 1024, 80, 0, fmax)
    assert mine.dtype == np.float32 and mine.shape == (80, 513)
    assert np.array_equal(mine, ref)


@pytest.mark.parametrize("cfg_key,cls", [("hifigan_v1", "HiFiGAN"), ("istftnet", "iSTFTNet"), ("hifigan_small_rb2", "HiFiGAN")])
def test_constructor_matches_reference_state_dict(cfg_key, cls):
    """Same keys, shapes AND values as the reference constructor under the same seed
    (tests/golden/state_*.json was produced by the reference's own __init__)."""
    cfg = synth.CONFIGS[cfg_key]
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", f"state_{cfg_key}.json")))["tensors"]
    torch.manual_seed(cfg["seed"])
    gen = getattr(pkg, cls)(synth.AttrDict(cfg))
    sd = gen.state_dict()
    assert list(sd.keys()) == list(golden.keys())
    for k, v in sd.items():
        assert list(v.shape) == golden[k]["shape"], k
        assert abs(float(v.double().sum()) - golden[k]["sum"]) <= 1e-9 * max(1.0, golden[k]["abs"]), k
        assert abs(float(v.double().abs().sum()) - golden[k]["abs"]) <= 1e-9 * max(1.0, golden[k]["abs"]), k


def test_param_count_matches_published():
    gen = pkg.HiFiGAN(synth.AttrDict(synth.HIFIGAN_V1))
    n = sum(p.numel() for p in gen.parameters())
    assert round(n / 1e6, 1) == 13.9  # figure/Results_comparisons.png Table III
    gen = pkg.iSTFTNet(synth.AttrDict(synth.ISTFTNET))
    assert round(sum(p.numel() for p in gen.parameters()) / 1e6, 1) == 13.3


def test_remove_weight_norm_and_checkpoint_roundtrip(tmp_path, capsys):
    cfg = synth.HIFIGAN_SMALL
    gen = pkg.HiFiGAN(synth.AttrDict(cfg))
    state = synth.make_state(cfg, 3, "unit")
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    path = tmp_path / "g_00000001"
    torch.save({"generator": gen.state_dict()}, path)        # train_time_wi_inv.py:252-255 format
    gen2 = pkg.HiFiGAN(synth.AttrDict(cfg))
    gen2.load_state_dict(torch.load(path, map_location="cpu")["generator"])
    for (k1, v1), (k2, v2) in zip(gen.state_dict().items(), gen2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    gen2.remove_weight_norm()
    assert "Removing weight norm" in capsys.readouterr().out  # hifigan.py:127
    keys = set(gen2.state_dict().keys())
    assert "conv_pre.weight" in keys and not any(k.endswith(("weight_g", "weight_v")) for k in keys)
    w = gen2.state_dict()["resblocks.2.convs1.1.weight"].numpy()
    ref = np_oracle.weight_norm_fold(state["resblocks.2.convs1.1.weight_v"], state["resblocks.2.convs1.1.weight_g"])
    assert np.abs(w - ref).max() < 1e-6
    print(gen2)  # train_time_wi_inv.py:59 prints the module


def test_make_config_reads_cfg_json_fields():
    cfg = pkg._engine.make_config(synth.AttrDict(synth.ISTFTNET), lib_mod.GEN_ISTFTNET)
    assert (cfg.kind, cfg.in_channels, cfg.initial_channel, cfg.num_upsamples) == (1, 80, 512, 2)
    assert list(cfg.upsample_rates)[:2] == [8, 8] and list(cfg.upsample_kernel_sizes)[:2] == [16, 16]
    assert list(cfg.resblock_kernel_sizes)[:3] == [3, 7, 11] and list(cfg.resblock_dilations[2])[:3] == [1, 3, 5]
    assert (cfg.istft_n_fft, cfg.istft_hop) == (16, 4)
    cfg2 = pkg._engine.make_config(synth.AttrDict(synth.HIFIGAN_SMALL_RB2), lib_mod.GEN_HIFIGAN)
    assert cfg2.resblock_type == 2 and list(cfg2.num_dilations)[:3] == [2, 2, 2]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    gen = pkg.HiFiGAN(synth.AttrDict(synth.HIFIGAN_SMALL)).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        gen(torch.zeros(1, 80, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.mel_spectrogram(torch.zeros(1, 4096), 1024, 80, 22050, 256, 1024, 0, 8000)


def test_training_mode_has_no_cpu_fallback_either():
    """Train mode takes the CUDA forward-with-tape / backward path (never a torch fallback)."""
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    for gen in (pkg.HiFiGAN(synth.AttrDict(synth.HIFIGAN_SMALL)), pkg.iSTFTNet(synth.AttrDict(synth.ISTFTNET_SMALL))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            gen.train()(torch.zeros(1, 80, 4))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import it."""
    src_dir = os.path.join(ROOT, "neural-vocoders-as-speech-enhancers_b200")
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), (dirpath, f)
                assert "np_oracle" not in text and "torch_port" not in text, (dirpath, f)
