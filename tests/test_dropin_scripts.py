"""The drop-in promise of SURVEY.md 8(b) / 7.2: the reference's scripts run BYTE-IDENTICAL, only the environment
changes (``PYTHONPATH=<package>/dropin``).

  * CPU tests: the names the scripts import resolve to the drop-in (also when the script's own directory, which holds
    the reference's dataset.py / Models/, is first on sys.path), everything off the accelerated path stays the
    reference's own code, and the ``in_dataset=True`` data-loader call runs the reference's CPU function.
  * GPU tests: ``infers/inference_hifigan.py``, ``infers/inference_istftnet.py`` and ``train_time_wi_inv.py`` are run
    unmodified twice in subprocesses -- once against the reference's own modules (CPU for inference, stock PyTorch
    fp32 on the GPU for the trainer), once with the drop-in on PYTHONPATH -- on a synthetic wav list and a
    reference-format checkpoint, and their outputs (PCM_16 wav files, the generator checkpoint after two optimiser
    steps) are compared.

The scripts come from ``baseline/_ref`` (a plain copy of the reference files made by ``__graft_entry__.build()`` in the
build container; git-ignored, it travels to the GPU box).  librosa / soundfile / matplotlib are not in this image:
``tests/shims`` supplies test-harness stand-ins."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
PKG = os.path.join(ROOT, "neural-vocoders-as-speech-enhancers_b200")
DROPIN = os.path.join(PKG, "dropin")
SHIMS = os.path.join(ROOT, "tests", "shims")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train_time_wi_inv.py")),
                               reason="baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)")


def _env(dropin, extra=None):
    env = dict(os.environ)
    path = ([DROPIN] if dropin else []) + [SHIMS, ROOT]
    if env.get("PYTHONPATH"):
        path.append(env["PYTHONPATH"])  # e.g. the GPU box's own site hook
    env["PYTHONPATH"] = os.pathsep.join(path)
    env.setdefault("CUDA_MODULE_LOADING", "LAZY")
    env.update(extra or {})
    return env


def _run(args, cwd, env, timeout=900):
    p = subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, f"{args} failed ({p.returncode})\n--- stdout\n{p.stdout[-3000:]}\n--- stderr\n{p.stderr[-6000:]}"
    return p


def _write_wav(path, x, sr=22050):
    from scipy.io import wavfile
    wavfile.write(path, sr, np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16))


def _read_wav(path):
    from scipy.io import wavfile
    sr, d = wavfile.read(path)
    assert d.dtype == np.int16
    return sr, d


def _workdir(tmp_path, cfg_name, lengths, seed, regime="init", **overrides):
    """Synthetic corpus + reference-format checkpoint + a cfgs/*.json with its machine-specific paths replaced."""
    wavs = tmp_path / "wavs"
    wavs.mkdir()
    names = []
    for i, n in enumerate(lengths):
        # band-limited-ish signal so that PCM_16 quantisation of the input is benign: noise + two tones
        t = np.arange(n) / 22050.0
        x = 0.25 * synth.make_wave(1, n, seed + i)[0] + 0.2 * np.sin(2 * np.pi * (180 + 40 * i) * t) + 0.1 * np.sin(2 * np.pi * 1330 * t)
        _write_wav(str(wavs / f"utt{i}.wav"), x.astype(np.float32))
        names.append(f"utt{i}.wav")
    flist = tmp_path / "filelist.txt"
    flist.write_text("".join(f"DUMMY1/{n}|synthetic utterance\n" for n in names))
    cfg = json.load(open(os.path.join(REF, "cfgs", cfg_name)))
    model_cfg = synth.HIFIGAN_V1 if cfg["model_name"] == "HiFiGAN" else synth.ISTFTNET
    ckpt = tmp_path / "g_init"
    state = {k: torch.from_numpy(v) for k, v in synth.make_state(model_cfg, 1234, regime).items()}
    torch.save({"generator": state}, str(ckpt))
    cfg.update(input_training_wav_list=str(flist), input_validation_wav_list=str(flist), raw_wavfile_path=str(wavs),
               test_input_wavs_dir=str(flist), test_output_dir=str(tmp_path / "out"), checkpoint_file_load=str(ckpt),
               checkpoint_path=str(tmp_path / "ckpt"))
    cfg.update(overrides)
    cfg_path = tmp_path / cfg_name
    cfg_path.write_text(json.dumps(cfg))
    return cfg, str(cfg_path), names


# ---------------------------------------------------------------------------------------------------------
# CPU: name resolution
# ---------------------------------------------------------------------------------------------------------
_RESOLVE = r"""
import sys
sys.path.insert(0, {ref!r})              # what `python train_time_wi_inv.py` does: the script directory first
import torch
import dataset, Models
from dataset import Dataset, mel_spectrogram, get_dataset_filelist, load_wav, inverse_mel, amp_pha_specturm
from Models import HiFiGAN, iSTFTNet, HDDemucas, ConvTasNet
from Models.models import MultiPeriodDiscriminator, MultiScaleDiscriminator, feature_loss
from utils import AttrDict
assert dataset.__file__.startswith({dropin!r}), dataset.__file__
assert Models.__file__.startswith({dropin!r}), Models.__file__
assert Dataset.__module__ == "_nvse_reference_dataset" and Dataset.__init__.__code__.co_filename.startswith({ref!r})
assert load_wav.__code__.co_filename.startswith({ref!r})
assert mel_spectrogram.__code__.co_filename.startswith({dropin!r})
assert HiFiGAN.__module__.startswith("neural-vocoders-as-speech-enhancers_b200"), HiFiGAN.__module__
assert iSTFTNet.__module__.startswith("neural-vocoders-as-speech-enhancers_b200")
assert HDDemucas.__module__ == "Models.hddemucas" and sys.modules["Models.hddemucas"].__file__.startswith({ref!r})
import os
if os.environ.get("NVSE_B200_DISCRIMINATORS", "0") == "1":   # opt-in: B200-backed discriminators
    assert MultiPeriodDiscriminator.__module__.startswith("neural-vocoders-as-speech-enhancers_b200")
    assert MultiScaleDiscriminator.__module__.startswith("neural-vocoders-as-speech-enhancers_b200")
else:                                                        # default: the reference's own classes
    assert MultiPeriodDiscriminator.__module__ == "_nvse_reference_models" and MultiScaleDiscriminator.__module__ == "_nvse_reference_models"
import Models.models
assert Models.models.__file__.startswith({dropin!r})
assert feature_loss.__code__.co_filename.startswith({ref!r})  # everything else of Models/models.py stays the reference's own
assert Models.models.MultiResolutionDiscriminator.__module__ == "_nvse_reference_models"
import Models.hifigan, Models.istftnet
assert Models.hifigan.__file__.startswith({dropin!r}) and Models.hifigan.HiFiGAN is HiFiGAN
# the data-loader call (dataset.py:218-241): CPU tensors in, the reference's own function, no CUDA involved
y = (torch.rand(1, 4096) - 0.5)
m = mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000, center=True, in_dataset=True)
r = sys.modules["_nvse_reference_dataset"].mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000, center=True, in_dataset=True)
assert m.shape == (1, 80, 17) and torch.equal(m, r) and not m.is_cuda
# without a GPU the accelerated call fails loudly instead of falling back
if not torch.cuda.is_available():
    try:
        mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    except Exception as e:
        assert "CUDA" in str(e), e
    else:
        raise SystemExit("mel_spectrogram ran without a CUDA device")
h = AttrDict(dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=32,
                  resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, model_name="HiFiGAN"))
g = eval(h.model_name)(h)
assert sorted(g.state_dict())[:3] == ["conv_post.bias", "conv_post.weight_g", "conv_post.weight_v"]
print("RESOLVE-OK")
"""


@needs_ref
@pytest.mark.parametrize("b200_discriminators", ["0", "1"])
def test_names_resolve_to_dropin_even_behind_the_script_directory(b200_discriminators):
    p = _run(["-c", _RESOLVE.format(ref=REF, dropin=DROPIN)], cwd=REF, env=_env(True, {"NVSE_B200_DISCRIMINATORS": b200_discriminators}))
    assert "RESOLVE-OK" in p.stdout


@needs_ref
def test_without_dropin_the_reference_modules_load():
    code = ("import sys; sys.path.insert(0, %r); import dataset, Models; "
            "assert dataset.__file__.startswith(%r) and Models.HiFiGAN.__module__ == 'Models.hifigan'; print('REF-OK')" % (REF, REF))
    assert "REF-OK" in _run(["-c", code], cwd=REF, env=_env(False)).stdout


def test_sitecustomize_chains_to_a_later_site_hook(tmp_path):
    (tmp_path / "sitecustomize.py").write_text("import os; os.environ['NVSE_CHAINED_HOOK'] = '1'\n")
    env = _env(True)
    env["PYTHONPATH"] = env["PYTHONPATH"] + os.pathsep + str(tmp_path)
    p = _run(["-c", "import os, sys; print(os.environ.get('NVSE_CHAINED_HOOK'), any(type(f).__name__ == '_NvseDropinFinder' for f in sys.meta_path))"],
             cwd=str(tmp_path), env=env)
    assert p.stdout.split() == ["1", "True"]


# ---------------------------------------------------------------------------------------------------------
# GPU: the unmodified scripts
# ---------------------------------------------------------------------------------------------------------
def _snr_db(ref, deg):
    """Metrics/snr.py:25-31 (de-meaned)."""
    ref = ref.astype(np.float64) - ref.mean()
    deg = deg.astype(np.float64) - deg.mean()
    return 10 * np.log10((ref ** 2).sum() / max(((ref - deg) ** 2).sum(), 1e-30))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("script,cfg_name", [("inference_hifigan.py", "hifigan_v1_config.json"),
                                             ("inference_istftnet.py", "istftnet_config.json")])
def test_inference_script_unmodified(tmp_path, script, cfg_name):
    from util import report
    # "unit" weights: a full-scale output waveform, so that the PCM_16 quantisation the script applies (1 LSB = 3e-5) is
    # far below the arithmetic differences being measured (at the reference's random init the output is a DC offset
    # plus ~100 LSB of signal and the SNR of ANY two PCM files is quantisation-limited to ~30 dB)
    cfg, cfg_path, names = _workdir(tmp_path, cfg_name, [22050, 15000, 30001], seed=300, regime="unit")
    outs = {}
    for arm, dropin, extra in (("reference", False, {"CUDA_VISIBLE_DEVICES": ""}),
                               ("dropin_bf16", True, {}),
                               ("dropin_fp32", True, {"NVSE_B200_PRECISION": "fp32"})):
        out_dir = tmp_path / f"out_{arm}"
        c = dict(cfg, test_output_dir=str(out_dir))
        cp = tmp_path / f"{arm}.json"
        cp.write_text(json.dumps(c))
        p = _run([os.path.join(REF, "infers", script), "--cfg_filename", str(cp)], cwd=os.path.join(REF, "infers"), env=_env(dropin, extra))
        outs[arm] = {n: _read_wav(str(out_dir / n)) for n in names}
        assert "Removing weight norm" in p.stdout
    for n in names:
        sr, ref = outs["reference"][n]
        assert sr == 22050 and len(ref) == 256 * (1 + (len(ref) - 1) // 256)
        _, f32 = outs["dropin_fp32"][n]
        _, b16 = outs["dropin_bf16"][n]
        assert f32.shape == ref.shape == b16.shape
        lsb = int(np.abs(f32.astype(np.int32) - ref.astype(np.int32)).max())
        snr = _snr_db(ref, b16)
        report(f"dropin {script} {n}: fp32 path max |PCM_16 diff| {lsb} LSB (<= 4 = 1e-4 + rounding), SNR {_snr_db(ref, f32):.1f} dB; "
               f"16-bit path SNR {snr:.1f} dB (>= 40); output rms {ref.astype(np.float64).std():.0f} LSB")
        assert lsb <= 4
        assert snr >= 40.0


@needs_ref
@pytest.mark.gpu
def test_train_script_unmodified(tmp_path):
    """Two optimiser steps of train_time_wi_inv.py:166-236 (num_workers = 2: the in_dataset=True mel runs in forked
    workers), drop-in (fp32 training path) vs the reference modules under stock PyTorch fp32 on the same GPU."""
    from util import report
    seg = 8192
    cfg, cfg_path, names = _workdir(tmp_path, "hifigan_v1_config.json", [seg] * 4 + [6000], seed=500, batch_size=2, segment_size=seg,
                                    num_workers=2, training_epochs=1, checkpoint_interval=1, summary_interval=1, stdout_interval=1,
                                    validation_interval=1000)
    # validation list: one short utterance (the validation pass runs at step 0 whatever the interval)
    vlist = tmp_path / "val.txt"
    vlist.write_text(f"DUMMY1/{names[-1]}|v\n")
    tlist = tmp_path / "train.txt"
    tlist.write_text("".join(f"DUMMY1/{n}|t\n" for n in names[:4]))
    ck = {}
    # NVSE_B200_DISCRIMINATORS=1: generator AND both discriminators on this repo's kernels (the drop-in's default keeps the
    # reference's discriminator classes, which run cuDNN's faster TF32 convolutions)
    for arm, dropin, extra in (("reference", False, {"NVIDIA_TF32_OVERRIDE": "0"}), ("dropin", True, {"NVSE_B200_DISCRIMINATORS": "1"})):
        c = dict(cfg, checkpoint_path=str(tmp_path / f"ckpt_{arm}"), input_training_wav_list=str(tlist), input_validation_wav_list=str(vlist))
        cp = tmp_path / f"train_{arm}.json"
        cp.write_text(json.dumps(c))
        _run([os.path.join(REF, "train_time_wi_inv.py"), "--cfg_filename", str(cp)], cwd=REF, env=_env(dropin, extra), timeout=1500)
        path = os.path.join(c["checkpoint_path"], "g_00000001")
        assert os.path.isfile(path), os.listdir(c["checkpoint_path"])
        ck[arm] = torch.load(path, map_location="cpu")["generator"]
    assert sorted(ck["reference"]) == sorted(ck["dropin"])
    # the initial weights both runs started from: the constructor under torch.manual_seed(h.seed) (train_time_wi_inv.py:46,52)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    torch.manual_seed(cfg["seed"])
    w0 = pkg.HiFiGAN(synth.AttrDict(cfg)).state_dict()
    num = den_r = den_d = 0.0
    worst = 0.0
    for k, wr in ck["reference"].items():
        ur, ud = (wr - w0[k]).double().flatten(), (ck["dropin"][k] - w0[k]).double().flatten()
        num += float(ur @ ud); den_r += float(ur @ ur); den_d += float(ud @ ud)
        worst = max(worst, float((wr - ck["dropin"][k]).abs().max()))
    cos = num / max((den_r * den_d) ** 0.5, 1e-30)
    report(f"dropin train_time_wi_inv.py, 2 steps: cosine(update_reference, update_dropin) {cos:.5f} (>= 0.98), "
           f"|update| ref {den_r ** 0.5:.3e} / dropin {den_d ** 0.5:.3e}, worst per-weight difference {worst:.2e} (one AdamW step = 2e-4)")
    assert den_r > 0 and den_d > 0
    assert cos >= 0.98
