"""Pin both oracle restatements against outputs of the reference's own code
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import synth
from oracle import np_oracle, torch_port

A = synth.HIFIGAN_V1
MEL_CASES = ["mel_b2_t4100", "mel_b1_t513", "mel_b3_t8192_fmax_half", "mel_1d_t22050", "mel_tone_silence"]
GEN_CASES = ["hifigan_v1_init_f6", "hifigan_v1_unit_f9", "hifigan_small_unit_f33", "hifigan_small_f1",
             "hifigan_small_rb2_f17", "istftnet_init_f5", "istftnet_unit_f7", "istftnet_small_unit_f40"]


@pytest.mark.parametrize("name", MEL_CASES)
def test_mel_np_oracle_matches_reference(name):
    g = synth.load_golden(name)
    out = np_oracle.mel_spectrogram(g["y"], A["n_fft"], A["num_mels"], A["sampling_rate"], A["hop_size"],
                                    A["win_size"], A["fmin"], g["meta"]["fmax"])
    assert out.shape == g["out"].shape and out.dtype == np.float32
    # reference is fp32 (pocketfft + fp32 bmm); the float64 oracle differs by fp32 rounding only
    assert synth.mel_mismatch(out, g["out"]) <= 1.0
    assert np.abs(out - g["out"]).mean() <= 2e-5


@pytest.mark.parametrize("name", MEL_CASES)
def test_mel_torch_port_matches_reference(name):
    g = synth.load_golden(name)
    out = torch_port.mel_spectrogram(torch.from_numpy(g["y"]), A["n_fft"], A["num_mels"], A["sampling_rate"],
                                     A["hop_size"], A["win_size"], A["fmin"], g["meta"]["fmax"]).numpy()
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 1e-5  # same library calls -> same bits up to threading


def test_clamp_floor_is_hit():
    g = synth.load_golden("mel_tone_silence")
    assert np.isclose(g["out"].min(), np.log(1e-5), atol=1e-6)


@pytest.mark.parametrize("name", GEN_CASES)
def test_generator_np_oracle_matches_reference(name):
    g = synth.load_golden(name)
    m = g["meta"]
    cfg = synth.CONFIGS[m["cfg"]]
    state = synth.make_state(cfg, m["weight_seed"], m["regime"])
    fwd = np_oracle.hifigan_forward if cfg["model_name"] == "HiFiGAN" else np_oracle.istftnet_forward
    out = fwd(state, cfg, g["mel"])
    assert out.shape == g["out"].shape
    scale = max(1e-3, np.abs(g["out"]).max())
    assert np.abs(out - g["out"]).max() <= 2e-5 * max(1.0, scale)


@pytest.mark.parametrize("name", GEN_CASES)
def test_generator_torch_port_matches_reference(name):
    g = synth.load_golden(name)
    m = g["meta"]
    cfg = synth.CONFIGS[m["cfg"]]
    folded = torch_port.fold_state(synth.make_state(cfg, m["weight_seed"], m["regime"]))
    fwd = torch_port.hifigan_forward if cfg["model_name"] == "HiFiGAN" else torch_port.istftnet_forward
    out = fwd(folded, cfg, torch.from_numpy(g["mel"])).numpy()
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 2e-5 * max(1.0, np.abs(g["out"]).max())


def test_istft_head_matches_reference():
    g = synth.load_golden("istft_head_t37")
    out = np_oracle.istft_head(g["mag"], g["phase"], 16, 4)
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= 5e-6 * np.abs(g["out"]).max()


def test_e2e_matches_reference():
    g = synth.load_golden("e2e_hifigan_small_t5000")
    m = g["meta"]
    cfg = synth.CONFIGS[m["cfg"]]
    mel = np_oracle.mel_spectrogram(g["y"], cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"],
                                    cfg["win_size"], cfg["fmin"], cfg["fmax"])
    assert synth.mel_mismatch(mel, g["mel"]) <= 1.0
    out = np_oracle.hifigan_forward(synth.make_state(cfg, m["weight_seed"], m["regime"]), cfg, mel)
    assert np.abs(out - g["out"]).max() <= 1e-4


def test_mel_filterbank_against_torchaudio():
    """librosa.filters.mel is not installed ("parity unpinned"): cross-check the restatement
    against torchaudio's independent Slaney implementation, and its published structure
    (727 non-zeros, 142 all-zero columns, max 0.026493)."""
    import torchaudio
    for fmax in (8000.0, 11025.0):
        w = np_oracle.mel_filterbank(22050, 1024, 80, 0, fmax)
        ta = torchaudio.functional.melscale_fbanks(513, 0.0, fmax, 80, 22050, norm="slaney", mel_scale="slaney").T.numpy()
        assert w.shape == (80, 513) and w.dtype == np.float32
        assert np.abs(w - ta).max() < 2e-7
    w = np_oracle.mel_filterbank(22050, 1024, 80, 0, 8000)
    assert int((w != 0).sum()) == 727
    assert int((np.abs(w).sum(0) == 0).sum()) == 142
    assert abs(float(w.max()) - 0.026493) < 1e-6


def test_snr_definition():
    rng = np.random.default_rng(0)
    ref = rng.normal(size=1000) + 5.0
    deg = ref + 0.01 * rng.normal(size=1000) + 0.3  # DC offset is removed by the de-meaned SNR
    assert np_oracle.snr_db(ref, deg) > 35
    assert np_oracle.snr_db(ref, deg, demean=False) < 30
