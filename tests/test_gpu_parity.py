"""Parity of the CUDA path (through the C ABI / the drop-in Python API) against the
oracle, the reference-generated golden vectors and fp32 PyTorch ops.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import np_oracle
from util import build_generator, conv1d_cl, conv_transpose1d_cl, lib_mod, pkg, report, resblock1_cl, stream_ptr

pytestmark = pytest.mark.gpu
A = synth.HIFIGAN_V1
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _mel(y, fmax=A["fmax"], **kw):
    return pkg.mel_spectrogram(y, A["n_fft"], A["num_mels"], A["sampling_rate"], A["hop_size"], A["win_size"],
                               A["fmin"], fmax, **kw)


# ------------------------------------------------------------------------------ front-end
@pytest.mark.parametrize("name", ["mel_b2_t4100", "mel_b1_t513", "mel_b3_t8192_fmax_half", "mel_1d_t22050",
                                  "mel_tone_silence"])
def test_mel_matches_reference_golden(name):
    g = synth.load_golden(name)
    out = _mel(torch.from_numpy(g["y"]).to(DEV), g["meta"]["fmax"])
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == g["out"].shape
    ratio = synth.mel_mismatch(out.cpu().numpy(), g["out"])
    assert ratio <= 1.0, ratio
    # north_star gate for the front-end: log-mel L1 <= 1e-3 (we are two orders below it even
    # on the tone fixture, whose near-floor bins are pure fp32 FFT round-off in the reference too)
    assert np.abs(out.cpu().numpy() - g["out"]).mean() <= 2e-5


@pytest.mark.parametrize("batch,samples,seed", [(1, 600, 1), (3, 1024, 2), (2, 1279, 3), (5, 2560, 4), (2, 22051, 5)])
def test_mel_matches_oracle_ragged_sizes(batch, samples, seed):
    y = synth.make_wave(batch, samples, seed)
    out = _mel(torch.from_numpy(y).to(DEV)).cpu().numpy()
    ref = np_oracle.mel_spectrogram(y, A["n_fft"], A["num_mels"], A["sampling_rate"], A["hop_size"], A["win_size"],
                                    A["fmin"], A["fmax"])
    assert out.shape == ref.shape == (batch, 80, 1 + samples // 256)
    assert synth.mel_mismatch(out, ref) <= 1.0


def test_mel_cpu_input_and_in_dataset_return_on_cpu():
    y = torch.from_numpy(synth.make_wave(2, 3000, 6))
    a = _mel(y)
    b = _mel(y.to(DEV), in_dataset=True)
    assert a.device.type == "cpu" and b.device.type == "cpu"
    assert torch.equal(a, b)
    assert any(k.endswith("-cpu") for k in pkg.dataset.mel_window)  # reference cache keys stay populated


def test_mel_strided_rows_and_short_input():
    big = torch.from_numpy(synth.make_wave(4, 5000, 7)).to(DEV)
    view = big[:, :4096]  # row stride 5000
    assert torch.equal(_mel(view), _mel(view.contiguous()))
    with pytest.raises(RuntimeError, match="reflect"):
        _mel(torch.zeros(1, 512, device=DEV))  # torch.stft also refuses pad >= T


def test_mel_full_size_cfg2_properties():
    """BASELINE cfg2 (64 x 4 s): against torch.stft on the GPU in fp32, plus a size-independent
    property: shifting the waveform by two hops shifts interior frames by two, bit-exactly."""
    y = torch.from_numpy(synth.make_wave(64, 88200, 8)).to(DEV)
    out = _mel(y)
    assert tuple(out.shape) == (64, 80, 345)
    basis = torch.from_numpy(np_oracle.mel_filterbank(22050, 1024, 80, 0, 8000)).to(DEV)
    spec = torch.stft(y, 1024, hop_length=256, win_length=1024, window=torch.hann_window(1024, device=DEV),
                      center=True, return_complex=True).abs()
    ref = torch.log(torch.clamp(torch.einsum("mk,bkf->bmf", basis, spec), min=1e-5))
    assert synth.mel_mismatch(out.cpu().numpy(), ref.cpu().numpy()) <= 1.0
    shifted = _mel(y[:, 512:])
    assert torch.equal(shifted[:, :, 2:300], out[:, :, 4:302])


# ------------------------------------------------------------------------------ layers
CONV_CASES = [  # (Cin, Cout, k, d, T, B): every (k, d) of the MRF + the edge layers
    (32, 32, 3, 1, 300, 2), (32, 32, 3, 3, 257, 1), (32, 32, 3, 5, 64, 2), (32, 32, 7, 1, 129, 2), (32, 32, 7, 3, 500, 1),
    (32, 32, 7, 5, 200, 1), (32, 32, 11, 1, 128, 1), (32, 32, 11, 3, 100, 2), (32, 32, 11, 5, 333, 1),
    (64, 64, 7, 3, 260, 2), (128, 128, 11, 5, 140, 1), (256, 256, 3, 1, 70, 1), (80, 512, 7, 1, 33, 2),
    (32, 1, 7, 1, 700, 2), (128, 18, 7, 1, 513, 2), (64, 64, 3, 1, 1, 1), (64, 64, 11, 5, 7, 1),
]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("cin,cout,k,d,T,B", CONV_CASES)
def test_conv1d_f32_matches_torch(cin, cout, k, d, T, B):
    x, w, b = _rand((B, cin, T), 1), _rand((cout, cin, k), 2, 1.0 / np.sqrt(cin * k)), _rand((cout,), 3)
    ref = F.conv1d(F.leaky_relu(x, 0.1), w, b, dilation=d, padding=(k - 1) * d // 2)
    out = conv1d_cl(x, w, b, d, in_slope=0.1)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), float((out - ref).abs().max())


def test_conv1d_f32_fused_epilogue():
    """residual add + MRF 1/3 scaling + accumulate (hifigan.py:49,113-119) in one launch."""
    x, w, b = _rand((2, 64, 150), 4), _rand((64, 64, 7), 5, 0.05), _rand((64,), 6)
    res, y0 = _rand((2, 64, 150), 7), _rand((2, 64, 150), 8)
    ref = y0 + (F.conv1d(F.leaky_relu(x, 0.1), w, b, padding=3) + res) / 3
    out = conv1d_cl(x, w, b, 1, in_slope=0.1, residual_bct=res, out_scale=1.0 / 3, y0_bct=y0)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("cin,cout,k,u,T,B", [(512, 256, 16, 8, 9, 2), (256, 128, 16, 8, 40, 1), (128, 64, 4, 2, 130, 2),
                                             (64, 32, 4, 2, 257, 1), (32, 16, 8, 4, 5, 1), (64, 32, 16, 8, 1, 1)])
def test_conv_transpose1d_f32_matches_torch(cin, cout, k, u, T, B):
    x, w, b = _rand((B, cin, T), 9), _rand((cin, cout, k), 10, 1.0 / np.sqrt(cin * 2)), _rand((cout,), 11)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=u, padding=(k - u) // 2)
    out = conv_transpose1d_cl(x, w, b, u, (k - u) // 2, in_slope=0.1)
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), float((out - ref).abs().max())


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


TC_CONV_CASES = [  # every MRF (C, k, d) class of HiFi-GAN V1, plus ragged / tiny T
    (32, 32, 3, 1, 300, 2), (32, 32, 3, 3, 257, 1), (32, 32, 3, 5, 64, 2), (32, 32, 7, 1, 129, 2), (32, 32, 7, 3, 500, 1),
    (32, 32, 7, 5, 200, 1), (32, 32, 11, 1, 128, 1), (32, 32, 11, 3, 100, 2), (32, 32, 11, 5, 333, 1),
    (64, 64, 3, 1, 1, 1), (64, 64, 7, 3, 260, 2), (64, 64, 11, 5, 7, 1), (128, 128, 3, 3, 1000, 1),
    (128, 128, 11, 5, 140, 1), (256, 256, 3, 1, 70, 1), (256, 256, 7, 5, 129, 2), (256, 256, 11, 5, 400, 1),
]


@pytest.mark.parametrize("cin,cout,k,d,T,B", TC_CONV_CASES)
def test_conv1d_tensor_core_matches_torch_on_bf16_operands(cin, cout, k, d, T, B):
    """tcgen05 path: products of bf16-rounded operands are exact in fp32, so against an fp32 conv of
    the rounded operands only the accumulation order differs."""
    x, w, b = _rand((B, cin, T), 21), _rand((cout, cin, k), 22, 1.0 / np.sqrt(cin * k)), _rand((cout,), 23)
    res, y0 = _rand((B, cout, T), 24), _rand((B, cout, T), 25)
    conv = F.conv1d(_bf(F.leaky_relu(x, 0.1)), _bf(w), b, dilation=d, padding=(k - 1) * d // 2)
    out = conv1d_cl(x, w, b, d, in_slope=0.1, tc=True)
    assert not lib_mod.tc_abort_status()
    assert torch.allclose(out, conv, atol=3e-5, rtol=1e-5), float((out - conv).abs().max())
    out2 = conv1d_cl(x, w, b, d, in_slope=0.1, residual_bct=res, out_scale=1.0 / 3, y0_bct=y0, tc=True)
    assert torch.allclose(out2, y0 + (conv + res) / 3, atol=3e-5, rtol=1e-5)


RB_CASES = [  # (C, k, dilations, T, B): every MRF class of HiFi-GAN V1; T spans several CTA tiles, ragged ends, tiny T
    (32, 3, (1, 3, 5), 1300, 2), (32, 7, (1, 3, 5), 1000, 1), (32, 11, (1, 3, 5), 1000, 2), (32, 11, (1, 3, 5), 7, 1),
    (64, 3, (1, 3, 5), 700, 1), (64, 7, (1, 3, 5), 1500, 1), (64, 11, (1, 3, 5), 900, 2), (64, 11, (1, 3, 5), 393, 1),
    (128, 3, (1,), 600, 2), (128, 7, (3,), 255, 1), (128, 11, (5,), 1000, 1), (128, 11, (1, 3, 5), 400, 1),
    (256, 3, (5,), 300, 1), (256, 11, (3,), 130, 2), (32, 3, (1, 1), 50, 1), (64, 5, (2,), 2000, 1),
]


@pytest.mark.parametrize("C_,k,dils,T,B", RB_CASES)
def test_fused_resblock1_tensor_core_matches_torch_on_bf16_operands(C_, k, dils, T, B):
    """The fused chain rounds exactly where this reference does (conv operands to bf16, fp32
    accumulate, fp32 residual stream); what differs is fp32 summation order, which can flip an
    occasional bf16 rounding of an intermediate -- hence a loose max-abs and a tight mean."""
    n = len(dils)
    x = _rand((B, C_, T), 41)
    w1 = [_rand((C_, C_, k), 42 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    w2 = [_rand((C_, C_, k), 52 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    b1 = [_rand((C_,), 62 + m, 0.3) for m in range(n)]
    b2 = [_rand((C_,), 72 + m, 0.3) for m in range(n)]
    ref, ref32 = x, x
    for m, d in enumerate(dils):
        h = F.conv1d(_bf(F.leaky_relu(ref, 0.1)), _bf(w1[m]), b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref = F.conv1d(_bf(F.leaky_relu(h, 0.1)), _bf(w2[m]), b2[m], padding=(k - 1) // 2) + ref
        h = F.conv1d(F.leaky_relu(ref32, 0.1), w1[m], b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref32 = F.conv1d(F.leaky_relu(h, 0.1), w2[m], b2[m], padding=(k - 1) // 2) + ref32
    out = resblock1_cl(x, w1, b1, w2, b2, dils)
    assert not lib_mod.tc_abort_status()
    assert torch.isfinite(out).all()
    err = (out - ref).abs()
    # a flipped rounding is one bf16 ulp of one operand and spreads through the later convs (the same
    # reference evaluated by torch on CPU and on GPU differs by as much): bound it by a fraction of
    # what bf16 rounding itself costs against the unrounded fp32 chain
    budget = float((ref - ref32).abs().mean())
    assert float(err.max()) <= 2e-2 and float(err.mean()) <= 0.5 * budget + 1e-5, (float(err.max()), float(err.mean()), budget)
    y0 = _rand((B, C_, T), 99)
    out2 = resblock1_cl(x, w1, b1, w2, b2, dils, out_scale=1.0 / 3, y0_bct=y0)
    assert torch.allclose(out2, y0 + out / 3, atol=1e-5, rtol=1e-5)


RB_T32_CASES = [  # the T32-layout kernels the generator launches: pipelined pair kernel (C = 128) and the chain kernel
    (128, 3, (1,), 1000, 2), (128, 7, (3,), 700, 1), (128, 11, (5,), 1500, 2), (128, 11, (1,), 255, 1), (128, 3, (5,), 31, 1),
    (128, 7, (1,), 5000, 3), (128, 7, (3,), 55168, 2), (64, 7, (1, 3, 5), 1000, 1), (32, 11, (1, 3, 5), 900, 2), (256, 7, (3,), 300, 1),
]


@pytest.mark.parametrize("pairpipe", ["1", "0"])
@pytest.mark.parametrize("C_,k,dils,T,B", RB_T32_CASES)
def test_fused_resblock1_t32_layout_kernels(monkeypatch, C_, k, dils, T, B, pairpipe):
    """Same reference as above, through the T32 activation layout (and, for single C = 128 pairs, the
    persistent software-pipelined pair kernel): NVSE_RB_LAYER_T32 makes the layer entry convert to T32."""
    if pairpipe == "0" and not (C_ == 128 and len(dils) == 1):
        pytest.skip("only C = 128 single pairs have two kernels")
    monkeypatch.setenv("NVSE_RB_LAYER_T32", "1")
    monkeypatch.setenv("NVSE_PAIRPIPE", pairpipe)
    n = len(dils)
    x = _rand((B, C_, T), 141)
    w1 = [_rand((C_, C_, k), 142 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    w2 = [_rand((C_, C_, k), 152 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    b1 = [_rand((C_,), 162 + m, 0.3) for m in range(n)]
    b2 = [_rand((C_,), 172 + m, 0.3) for m in range(n)]
    ref, ref32 = x, x
    for m, d in enumerate(dils):
        h = F.conv1d(_bf(F.leaky_relu(ref, 0.1)), _bf(w1[m]), b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref = F.conv1d(_bf(F.leaky_relu(h, 0.1)), _bf(w2[m]), b2[m], padding=(k - 1) // 2) + ref
        h = F.conv1d(F.leaky_relu(ref32, 0.1), w1[m], b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref32 = F.conv1d(F.leaky_relu(h, 0.1), w2[m], b2[m], padding=(k - 1) // 2) + ref32
    out = resblock1_cl(x, w1, b1, w2, b2, dils)
    assert not lib_mod.tc_abort_status()
    assert torch.isfinite(out).all()
    err = (out - ref).abs()
    budget = float((ref - ref32).abs().mean())
    assert float(err.max()) <= 2e-2 and float(err.mean()) <= 0.5 * budget + 1e-5, (float(err.max()), float(err.mean()), budget)
    y0 = _rand((B, C_, T), 199)
    out2 = resblock1_cl(x, w1, b1, w2, b2, dils, out_scale=1.0 / 3, y0_bct=y0)
    assert torch.allclose(out2, y0 + out / 3, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("C_,k,dils,T,B", [(32, 3, (1, 3, 5), 1300, 2), (32, 11, (1, 3, 5), 1000, 1), (32, 7, (1, 3, 5), 40, 1),
                                           (64, 7, (1, 3), 700, 1)])
def test_fused_resblock1_half_precision_intermediate(monkeypatch, C_, k, dils, T, B):
    """The variant the generator runs at C <= 32: the c1 -> c2 intermediate and the c2 weights are rounded to IEEE
    half (fp16 tcgen05 operands) instead of bf16; everything else as above."""
    monkeypatch.setenv("NVSE_RB_H16", "1")
    f16 = lambda v: v.to(torch.float16).to(torch.float32)
    n = len(dils)
    x = _rand((B, C_, T), 241)
    w1 = [_rand((C_, C_, k), 242 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    w2 = [_rand((C_, C_, k), 252 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
    b1 = [_rand((C_,), 262 + m, 0.3) for m in range(n)]
    b2 = [_rand((C_,), 272 + m, 0.3) for m in range(n)]
    ref, ref32 = x, x
    for m, d in enumerate(dils):
        h = F.conv1d(_bf(F.leaky_relu(ref, 0.1)), _bf(w1[m]), b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref = F.conv1d(f16(F.leaky_relu(h, 0.1)), f16(w2[m]), b2[m], padding=(k - 1) // 2) + ref
        h = F.conv1d(F.leaky_relu(ref32, 0.1), w1[m], b1[m], dilation=d, padding=(k - 1) * d // 2)
        ref32 = F.conv1d(F.leaky_relu(h, 0.1), w2[m], b2[m], padding=(k - 1) // 2) + ref32
    out = resblock1_cl(x, w1, b1, w2, b2, dils)
    assert not lib_mod.tc_abort_status()
    err = (out - ref).abs()
    budget = float((ref - ref32).abs().mean())
    assert float(err.max()) <= 2e-2 and float(err.mean()) <= 0.5 * budget + 1e-5, (float(err.max()), float(err.mean()), budget)
    # and it is the more accurate variant: closer to the unrounded chain than the all-bf16 kernel
    monkeypatch.setenv("NVSE_RB_H16", "0")
    out_bf = resblock1_cl(x, w1, b1, w2, b2, dils)
    assert float((out - ref32).abs().mean()) < float((out_bf - ref32).abs().mean())


@pytest.mark.parametrize("cin,cout,k,u,T,B", [(512, 256, 16, 8, 9, 2), (256, 128, 16, 8, 140, 1), (128, 64, 4, 2, 130, 2),
                                             (64, 32, 4, 2, 257, 1), (64, 32, 16, 8, 1, 1)])
def test_conv_transpose1d_tensor_core_matches_torch_on_bf16_operands(cin, cout, k, u, T, B):
    x, w, b = _rand((B, cin, T), 26), _rand((cin, cout, k), 27, 1.0 / np.sqrt(cin * 2)), _rand((cout,), 28)
    ref = F.conv_transpose1d(_bf(F.leaky_relu(x, 0.1)), _bf(w), b, stride=u, padding=(k - u) // 2)
    out = conv_transpose1d_cl(x, w, b, u, (k - u) // 2, in_slope=0.1, tc=True)
    assert not lib_mod.tc_abort_status()
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, atol=3e-5, rtol=1e-5), float((out - ref).abs().max())


@pytest.mark.parametrize("cin,cout,k,u,T,B", [(512, 256, 16, 8, 9, 2), (256, 128, 16, 8, 140, 1), (128, 64, 4, 2, 130, 2),
                                             (64, 32, 4, 2, 257, 1)])
def test_conv_transpose1d_tensor_core_half_operands(monkeypatch, cin, cout, k, u, T, B):
    """The upsamplers of the generator run with IEEE-half operands (NVSE_TC_F16): exact products of the
    half-rounded operands, fp32 accumulation."""
    monkeypatch.setenv("NVSE_TC_F16", "1")
    f16 = lambda v: v.to(torch.float16).to(torch.float32)
    x, w, b = _rand((B, cin, T), 26), _rand((cin, cout, k), 27, 1.0 / np.sqrt(cin * 2)), _rand((cout,), 28)
    ref = F.conv_transpose1d(f16(F.leaky_relu(x, 0.1)), f16(w), b, stride=u, padding=(k - u) // 2)
    out = conv_transpose1d_cl(x, w, b, u, (k - u) // 2, in_slope=0.1, tc=True)
    assert not lib_mod.tc_abort_status()
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, atol=3e-5, rtol=1e-5), float((out - ref).abs().max())


def test_weight_norm_fold_both_layouts():
    lib = lib_mod.load()
    for shape in [(256, 256, 11), (512, 256, 16), (1, 32, 7)]:  # Conv1d [Cout,Cin,k]; ConvT [Cin,Cout,k]: dim 0 either way
        v, g = _rand(shape, 12, 0.01), _rand((shape[0], 1, 1), 13).abs() + 0.1
        w = torch.empty_like(v)
        lib_mod.check(lib.nvse_weight_norm_fold_f32(lib_mod.ptr(v), lib_mod.ptr(g), lib_mod.ptr(w), shape[0],
                                                    shape[1] * shape[2], stream_ptr()))
        ref = np_oracle.weight_norm_fold(v.cpu().numpy(), g.cpu().numpy())
        assert np.abs(w.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_istft_head_matches_reference_golden_and_edges():
    g = synth.load_golden("istft_head_t37")
    lib = lib_mod.load()
    # the head takes conv_post output z with mag = exp(z[:9]), phase = sin(z[9:]); invert those maps for the fixture
    z = np.concatenate([np.log(g["mag"]), np.arcsin(np.clip(g["phase"], -1, 1))], axis=1)  # [B, 18, Tp]
    zt = torch.from_numpy(z.transpose(0, 2, 1).copy()).float().to(DEV)
    B, Tp = zt.shape[0], zt.shape[1]
    out = torch.empty((B, 4 * (Tp - 1)), device=DEV)
    lib_mod.check(lib.nvse_istft_head_f32(lib_mod.ptr(zt), lib_mod.ptr(out), B, Tp, 16, 4, stream_ptr()))
    scale = np.abs(g["out"]).max()
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 2e-5 * scale
    ref = np_oracle.istft_head(np.exp(z[:, :9]), np.sin(z[:, 9:]), 16, 4)
    assert np.abs(out.cpu().numpy()[:, :8] - ref[:, :8]).max() <= 2e-5 * scale   # first samples: envelope != 1.5
    assert np.abs(out.cpu().numpy()[:, -8:] - ref[:, -8:]).max() <= 2e-5 * scale


# ------------------------------------------------------------------------------ generators
GEN_CASES = ["hifigan_v1_init_f6", "hifigan_v1_unit_f9", "hifigan_small_unit_f33", "hifigan_small_f1",
             "hifigan_small_rb2_f17", "istftnet_init_f5", "istftnet_unit_f7", "istftnet_small_unit_f40"]


def _golden_generator(name, remove_wn):
    g = synth.load_golden(name)
    m = g["meta"]
    cfg = synth.CONFIGS[m["cfg"]]
    gen = build_generator(cfg, synth.make_state(cfg, m["weight_seed"], m["regime"]), DEV, remove_wn)
    return g, gen


@pytest.mark.parametrize("name", GEN_CASES)
@pytest.mark.parametrize("remove_wn", [False, True])
def test_generator_fp32_matches_reference_golden(name, remove_wn):
    """north_star fp32 gate: max-abs waveform error <= 1e-4 vs the reference on the same inputs."""
    g, gen = _golden_generator(name, remove_wn)
    gen.precision = "fp32"
    with torch.no_grad():
        out = gen(torch.from_numpy(g["mel"]).to(DEV))
    assert out.is_cuda and tuple(out.shape) == g["out"].shape
    err = np.abs(out.cpu().numpy() - g["out"]).max()
    report(f"{name} (weight_norm {'removed' if remove_wn else 'kept'}): fp32 max-abs {err:.2e}")
    assert err <= 1e-4, err


@pytest.mark.parametrize("name", GEN_CASES)
def test_generator_bf16_matches_reference_golden(name):
    """north_star bf16 gate: waveform SNR >= 40 dB (Metrics/snr.py:25-31, de-meaned); raw SNR reported too."""
    g, gen = _golden_generator(name, True)
    gen.precision = "bf16"
    with torch.no_grad():
        out = gen(torch.from_numpy(g["mel"]).to(DEV)).cpu().numpy()
    snr, raw = np_oracle.snr_db(g["out"], out), np_oracle.snr_db(g["out"], out, demean=False)
    report(f"{name}: bf16 SNR de-meaned {snr:.1f} dB, raw {raw:.1f} dB, max-abs {np.abs(out - g['out']).max():.2e}")
    assert snr >= 40.0 and raw >= 40.0


def test_e2e_wav_to_wav_matches_reference_golden():
    """infers/inference_hifigan.py:82-84: x = get_mel(wav); y = generator(x)."""
    g = synth.load_golden("e2e_hifigan_small_t5000")
    cfg = synth.CONFIGS[g["meta"]["cfg"]]
    gen = build_generator(cfg, synth.make_state(cfg, g["meta"]["weight_seed"], g["meta"]["regime"]), DEV, True)
    gen.precision = "fp32"
    with torch.no_grad():
        mel = _mel(torch.from_numpy(g["y"]).to(DEV))
        out = gen(mel)
    assert synth.mel_mismatch(mel.cpu().numpy(), g["mel"]) <= 1.0
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-4


def test_generator_cpu_tensors_are_staged_through_the_gpu():
    """The shipped inference script forces device = cpu (infers/inference_hifigan.py:129)."""
    g, gen = _golden_generator("hifigan_small_f1", True)
    gen = gen.to("cpu")
    gen.precision = "fp32"
    with torch.no_grad():
        out = gen(torch.from_numpy(g["mel"]))
    assert out.device.type == "cpu"
    assert np.abs(out.numpy() - g["out"]).max() <= 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_generator_batch_items_are_independent(precision):
    """Per-utterance sharding relies on this: an utterance vocoded alone, or on
    another rank, gives bit-identical samples to the same utterance inside a batch."""
    cfg = synth.HIFIGAN_SMALL
    gen = build_generator(cfg, synth.make_state(cfg, 17, "unit"), DEV, True)
    gen.precision = precision
    mel = torch.from_numpy(synth.make_mel(5, 37, 18)).to(DEV)
    with torch.no_grad():
        full = gen(mel)
        parts = [gen(mel[i:i + 1]) for i in (0, 3, 4)]
        two = gen(mel[pkg.shard_range(5, 2, 1).start:])
    for p, i in zip(parts, (0, 3, 4)):
        assert torch.equal(p[0], full[i])
    assert torch.equal(two, full[pkg.shard_range(5, 2, 1).start:])


def test_generator_weight_update_is_picked_up():
    cfg = synth.HIFIGAN_SMALL
    gen = build_generator(cfg, synth.make_state(cfg, 19, "unit"), DEV, False)
    gen.precision = "fp32"
    mel = torch.from_numpy(synth.make_mel(1, 8, 20)).to(DEV)
    with torch.no_grad():
        a = gen(mel)
        gen.conv_post.bias.add_(0.25)
        b = gen(mel)
        gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state(cfg, 19, "unit").items()})
        c = gen(mel)
    assert not torch.equal(a, b) and torch.equal(a, c)


def test_full_size_cfg3_fp32_vs_bf16_and_torch():
    """BASELINE cfg3 shape at reduced batch (4 x 8 s, F = 690): fp32 path vs PyTorch fp32 ops on the
    GPU (max-abs <= 1e-4) and bf16 path vs fp32 path (SNR >= 40 dB), reference-init-like weights."""
    from oracle import torch_port
    cfg = synth.HIFIGAN_V1
    state = synth.make_state(cfg, 1234, "init")
    gen = build_generator(cfg, state, DEV, True)
    mel = torch.from_numpy(synth.make_mel(4, 690, 30)).to(DEV)
    folded = {k: v.to(DEV) for k, v in torch_port.fold_state(state).items()}
    with torch.no_grad():
        ref = torch_port.hifigan_forward(folded, cfg, mel)
        gen.precision = "fp32"
        o32 = gen(mel)
        gen.precision = "bf16"
        o16 = gen(mel)
    assert tuple(o32.shape) == (4, 176640)
    assert float((o32 - ref).abs().max()) <= 1e-4
    snr = np_oracle.snr_db(ref.cpu().numpy(), o16.cpu().numpy())
    report(f"cfg3-shape (4 x 690 frames, init weights seed 1234): fp32 max-abs {float((o32 - ref).abs().max()):.2e}; "
           f"bf16 SNR {snr:.1f} dB (de-meaned), raw {np_oracle.snr_db(ref.cpu().numpy(), o16.cpu().numpy(), False):.1f} dB")
    assert snr >= 40.0


@pytest.mark.parametrize("name,B,F", [("cfg3 at its real batch", 32, 690), ("cfg5 micro-batch (the benchmarked shape)", 32, 862),
                                      ("cfg1", 1, 173), ("training segment", 16, 65)])
def test_benchmark_shapes_wav_to_wav(name, B, F):
    """The shapes BASELINE.json quotes, end to end through the public API (wav -> mel_spectrogram -> HiFiGAN -> wav):
    mel vs torch.stft-based fp32 ops, fp32 path vs PyTorch fp32 ops on the GPU (max-abs <= 1e-4), 16-bit tensor-core
    path vs the same (de-meaned SNR >= 40 dB, Metrics/snr.py:25-31), per utterance."""
    from oracle import torch_port
    cfg = synth.HIFIGAN_V1
    state = synth.make_state(cfg, 1234, "init")
    gen = build_generator(cfg, state, DEV, True)
    T = (F - 1) * 256 + 17   # F = 1 + T // 256 frames
    wav = torch.from_numpy(synth.make_wave(B, T, 77)).to(DEV)
    folded = {k: v.to(DEV) for k, v in torch_port.fold_state(state).items()}
    basis = torch.from_numpy(np_oracle.mel_filterbank(cfg["sampling_rate"], cfg["n_fft"], cfg["num_mels"], cfg["fmin"], cfg["fmax"])).to(DEV)
    with torch.no_grad():
        spec = torch.stft(wav, cfg["n_fft"], hop_length=cfg["hop_size"], win_length=cfg["win_size"],
                          window=torch.hann_window(cfg["win_size"], device=DEV), center=True, return_complex=True)
        mel_ref = torch.log(torch.clamp(basis @ spec.abs(), min=1e-5))
        del spec
        mel = _mel(wav)
        assert tuple(mel.shape) == (B, 80, F)
        mm = synth.mel_mismatch(mel.cpu().numpy(), mel_ref.cpu().numpy(), rtol=1e-4, atol=2e-6)
        ref = torch_port.hifigan_forward(folded, cfg, mel)
        gen.precision = "fp32"
        e32 = float((gen(mel) - ref).abs().max())
        gen.precision = "bf16"
        o16 = gen(mel)
    assert tuple(o16.shape) == (B, 256 * F)
    r, o = ref.double(), o16.double()
    r, o = r - r.mean(-1, keepdim=True), o - o.mean(-1, keepdim=True)
    snr = float((10 * torch.log10(r.pow(2).sum(-1) / (r - o).pow(2).sum(-1))).min())
    report(f"{name} ({B} x {F} frames, wav -> wav): mel mismatch ratio {mm:.3f} (<=1 at rtol 1e-4), fp32 max-abs {e32:.2e} (<=1e-4), "
           f"16-bit path worst-utterance SNR {snr:.1f} dB (>=40)")
    assert mm <= 1.0
    assert e32 <= 1e-4
    assert snr >= 40.0
    assert not lib_mod.tc_abort_status()


@pytest.mark.parametrize("seed", [1234, 7, 99])
def test_full_size_cfg4_istftnet_fp32_vs_bf16_and_torch(seed):
    """BASELINE cfg4 shape at reduced batch (iSTFTNet, 4 x 8 s, F = 690) and the seeds SURVEY 8(d) names:
    fp32 path vs PyTorch fp32 ops (max-abs <= 1e-4), 16-bit tensor-core path SNR >= 40 dB."""
    from oracle import torch_port
    cfg = synth.ISTFTNET
    state = synth.make_state(cfg, seed, "init")
    gen = build_generator(cfg, state, DEV, True)
    mel = torch.from_numpy(synth.make_mel(4, 690, 30 + seed % 5)).to(DEV)
    folded = {k: v.to(DEV) for k, v in torch_port.fold_state(state).items()}
    with torch.no_grad():
        ref = torch_port.istftnet_forward(folded, cfg, mel)
        gen.precision = "fp32"
        o32 = gen(mel)
        gen.precision = "bf16"
        o16 = gen(mel)
    assert o32.reshape(4, -1).shape == (4, 176640)
    assert float((o32.reshape(4, -1) - ref.reshape(4, -1)).abs().max()) <= 1e-4
    snr = np_oracle.snr_db(ref.reshape(4, -1).cpu().numpy(), o16.reshape(4, -1).cpu().numpy())
    report(f"cfg4-shape iSTFTNet (4 x 690 frames, init weights seed {seed}): fp32 max-abs "
           f"{float((o32.reshape(4, -1) - ref.reshape(4, -1)).abs().max()):.2e}; 16-bit path SNR {snr:.1f} dB (de-meaned)")
    assert snr >= 40.0


@pytest.mark.parametrize("seed", [7, 99])
def test_full_size_cfg3_other_seeds_bf16_snr(seed):
    """cfg3 shape, the other weight seeds of SURVEY 8(d) (99 was the worst bf16 case of the survey's probe)."""
    from oracle import torch_port
    cfg = synth.HIFIGAN_V1
    state = synth.make_state(cfg, seed, "init")
    gen = build_generator(cfg, state, DEV, True)
    mel = torch.from_numpy(synth.make_mel(2, 690, 31)).to(DEV)
    folded = {k: v.to(DEV) for k, v in torch_port.fold_state(state).items()}
    with torch.no_grad():
        ref = torch_port.hifigan_forward(folded, cfg, mel)
        gen.precision = "bf16"
        o16 = gen(mel)
    snr = np_oracle.snr_db(ref.cpu().numpy(), o16.cpu().numpy())
    report(f"cfg3-shape (2 x 690 frames, init weights seed {seed}): 16-bit path SNR {snr:.1f} dB (de-meaned)")
    assert snr >= 40.0


def test_vocoder_run_list_ragged_matches_single_utterance_runs():
    """Ragged utterances (and the empty list): grouped by exact length, each result identical to
    vocoding that utterance alone, as the reference's per-file loop does."""
    cfg = synth.HIFIGAN_V1
    gen = build_generator(cfg, synth.make_state(cfg, 5, "init"), DEV, remove_wn=True)
    gen.precision = "bf16"
    voc = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=2, device=DEV)
    assert voc.run_list([]) == []
    lens = [3000, 1537, 3000, 2048, 1537, 3000]
    wavs = [torch.from_numpy(synth.make_wave(1, n, 40 + i)[0]) for i, n in enumerate(lens)]
    outs = voc.run_list(wavs)
    assert [o.numel() for o in outs] == [(1 + n // cfg["hop_size"]) * cfg["hop_size"] for n in lens]
    for w, o in zip(wavs, outs):
        alone = voc.run_list([w])[0]
        assert torch.equal(o, alone)


@pytest.mark.parametrize("cfg_key", ["hifigan_v1", "istftnet"])
def test_padded_ragged_batches_are_bit_identical_to_single_utterance_runs(cfg_key):
    """Utterances of DIFFERENT lengths in one zero-padded batch with per-utterance lengths (nvse_frontend_mel_ragged_f32,
    nvse_generator_forward_ragged): the front-end reflects each utterance at its own end and every generator kernel masks
    it at its own length, so each result equals -- bit for bit -- vocoding the utterance alone (the reference's per-file
    loop, infers/inference_hifigan.py:67-95).  Lengths chosen to end inside, on and just past tile / block boundaries of
    the stages; one batch mixes a 0.4 s and a 1.6 s utterance (max_pad=0.9).  iSTFTNet: the reflection-padded conv_post and
    the iSTFT head's overlap-add stop at every utterance's own last frame."""
    cfg = synth.CONFIGS[cfg_key]
    gen = build_generator(cfg, synth.make_state(cfg, 5, "init"), DEV, remove_wn=True)
    gen.precision = "bf16"
    h = synth.AttrDict(cfg)
    voc = pkg.Vocoder(gen, h, micro_batch=8, device=DEV)
    lens = [9000, 11000, 10500, 22050, 21000, 35000, 8192, 8191, 8193, 33024, 4097, 30000]
    wavs = [torch.from_numpy(synth.make_wave(1, n, 70 + i)[0]) for i, n in enumerate(lens)]
    alone = []
    for w in wavs:
        m = _mel(w.to(DEV)[None])
        with torch.no_grad():
            alone.append(gen(m).reshape(-1).cpu())
    for max_pad in (0.25, 0.9):
        outs = voc.run_list(wavs, max_pad=max_pad)
        assert [o.numel() for o in outs] == [(1 + n // cfg["hop_size"]) * cfg["hop_size"] for n in lens]
        for i, (o, a) in enumerate(zip(outs, alone)):
            assert torch.equal(o, a), (max_pad, i, lens[i], float((o - a).abs().max()))
    # the padded front-end by itself: frames of every utterance equal its single-utterance log-mel
    tmax = max(lens)
    batch = torch.zeros((len(lens), tmax), device=DEV)
    for j, w in enumerate(wavs):
        batch[j, :lens[j]] = w.to(DEV)
    mel = _mel(batch, lengths=torch.tensor(lens, dtype=torch.int32, device=DEV))
    for j, w in enumerate(wavs):
        f = 1 + lens[j] // cfg["hop_size"]
        assert torch.equal(mel[j, :, :f], _mel(w.to(DEV)[None])[0])
    assert not lib_mod.tc_abort_status()
    report(f"padded ragged batches, {cfg['model_name']} ({len(lens)} utterances, 4097 .. 35000 samples, up to 90 % padding): bit-identical to single-utterance runs")


@pytest.mark.parametrize("cfg_key", ["hifigan_v1", "istftnet"])
def test_fused_wav_to_wav_call_is_bit_identical_to_the_two_stage_path(cfg_key):
    """nvse_vocoder_forward (the front-end writes the log-mel straight into conv_pre's channels-last staging layout, then the
    generator: one library call, no [B, 80, F] tensor, no transpose pass) against mel_spectrogram + generator called one after
    the other: float, PCM_16 and ragged output, both generators."""
    cfg = synth.CONFIGS[cfg_key]
    gen = build_generator(cfg, synth.make_state(cfg, 8, "init"), DEV, remove_wn=True)
    gen.precision = "bf16"
    voc = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=3, device=DEV)
    assert voc.fused()
    wav = torch.from_numpy(synth.make_wave(5, 5000, 81)).to(DEV)
    l0 = lib_mod.launch_count()
    fused = voc.run_device(wav)
    n_fused = lib_mod.launch_count() - l0
    voc.no_fuse = True
    assert not voc.fused()
    l0 = lib_mod.launch_count()
    two = voc.run_device(wav)
    n_two = lib_mod.launch_count() - l0
    assert torch.equal(fused, two) and n_fused == n_two - 2   # one transpose launch fewer per micro-batch
    p_two = voc.vocode(wav[:2], pcm16=True)
    lens = torch.tensor([5000, 3100, 4097], dtype=torch.int32)
    r_two = voc.vocode(wav[:3], lengths=lens)
    voc.no_fuse = False
    assert torch.equal(voc.vocode(wav[:2], pcm16=True), p_two) and p_two.dtype == torch.int16
    r_fused = voc.vocode(wav[:3], lengths=lens)
    up = fused.shape[1] // (1 + 5000 // cfg["hop_size"])
    for b, n in enumerate(lens.tolist()):
        k = (1 + n // cfg["hop_size"]) * up
        assert torch.equal(r_fused[b, :k], r_two[b, :k])
    # a Vocoder built with an index-less device ("cuda" != "cuda:0" for torch) must not make the engine rebuild its handle
    h0 = gen._engine.handle.value
    voc2 = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=3, device="cuda")
    assert voc2.fused() and torch.equal(voc2.run_device(wav), fused) and gen._engine.handle.value == h0
    gen.precision = "fp32"     # the fp32 path has no fused call: the same entry point falls back to the two stages
    assert not voc.fused() and voc.run_device(wav[:1]).shape == fused[:1].shape
    assert not lib_mod.tc_abort_status()


def test_vocoder_run_host_overlapped_copies_match_device_run():
    """Host-buffer pipeline (copies on side streams, overlapped with the kernels): identical to the
    device-resident run, with an utterance count that is not a multiple of the micro-batch."""
    cfg = synth.HIFIGAN_V1
    gen = build_generator(cfg, synth.make_state(cfg, 6, "init"), DEV, remove_wn=True)
    gen.precision = "bf16"
    voc = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=2, device=DEV)
    wav = torch.from_numpy(synth.make_wave(5, 2816, 77)).pin_memory()
    ref = voc.run_device(wav.to(DEV)).cpu()
    for _ in range(2):  # second pass reuses the side streams
        out = voc.run_host(wav)
        torch.cuda.synchronize()
        assert tuple(out.shape) == (5, 12 * 256) and torch.equal(out.reshape(5, -1), ref.reshape(5, -1))


def test_pcm16_output_matches_libsndfile_rule():
    """sf.write(..., 'PCM_16') of the reference's inference script = lrint(x * 0x7FFF) (round half to even), clipped."""
    x = torch.cat([torch.linspace(-1.2, 1.2, 100001), torch.tensor([0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.5 / 32767, 0.0])]).to(DEV)
    cfg = synth.HIFIGAN_V1
    voc = pkg.Vocoder(None, synth.AttrDict(cfg), device=DEV)
    got = voc.pcm16(x).cpu().numpy()
    want = np.clip(np.rint(x.cpu().numpy().astype(np.float32) * np.float32(32767.0)), -32768, 32767).astype(np.int16)
    assert got.dtype == np.int16 and np.array_equal(got, want)
    assert voc.pcm16(torch.empty(0, device=DEV)).numel() == 0
    gen = build_generator(cfg, synth.make_state(cfg, 6, "init"), DEV, remove_wn=True)
    voc = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=2, device=DEV)
    wav = torch.from_numpy(synth.make_wave(3, 2816, 78)).pin_memory()
    f = voc.run_host(wav); torch.cuda.synchronize()
    q = voc.run_host(wav, pcm16=True); torch.cuda.synchronize()
    assert q.dtype == torch.int16 and np.array_equal(q.numpy(), np.rint(f.numpy() * np.float32(32767.0)).astype(np.int16))


@pytest.mark.parametrize("cfg_name,precision", [("hifigan_v1", "bf16"), ("hifigan_v1", "fp32"), ("hifigan_small", "bf16"), ("istftnet", "bf16")])
def test_forward_pcm16_is_the_quantised_forward(cfg_name, precision):
    """nvse_generator_forward_pcm16 (PCM_16 fused into conv_post on the tensor-core plan of HiFiGAN, a separate pass on the
    other plans) is bit-identical to quantising the float output of nvse_generator_forward (infers/inference_hifigan.py:89-95)."""
    cfg = synth.CONFIGS[cfg_name]
    gen = build_generator(cfg, synth.make_state(cfg, 8, "unit"), DEV, remove_wn=True)
    gen.precision = precision
    mel = torch.from_numpy(synth.make_mel(2, 37, 81)).to(DEV)
    with torch.no_grad():
        f = gen(mel)
        q = gen.forward_pcm16(mel)
    want = np.clip(np.rint(f.cpu().numpy() * np.float32(32767.0)), -32768, 32767).astype(np.int16)
    assert q.dtype == torch.int16 and tuple(q.shape) == tuple(f.shape)
    assert np.array_equal(q.cpu().numpy(), want)
    assert len(np.unique(want)) > 100  # a full-scale signal, not a constant
    with pytest.raises(RuntimeError, match="inference"):
        gen.train()
        gen._engine.forward(gen, mel.requires_grad_(True), pcm16=True)  # the training path has no PCM output


def test_generator_forward_is_deterministic_under_repetition():
    """The mbarrier protocols of the fused kernels carry no data race: 20 back-to-back forwards of the same
    batch (cfg3 frame count, all three fused kernel families) are bit-identical and trip no bounded wait."""
    cfg = synth.HIFIGAN_V1
    gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), DEV, remove_wn=True)
    gen.precision = "bf16"
    mel = torch.from_numpy(synth.make_mel(3, 690, 33)).to(DEV)
    with torch.no_grad():
        first = gen(mel).clone()
        for _ in range(20):
            assert torch.equal(gen(mel), first)
    assert not lib_mod.tc_abort_status()


@pytest.mark.parametrize("frames,batch", [(1, 1), (2, 3), (3, 2), (5, 1), (33, 2)])
def test_tiny_inputs_through_the_fused_t32_path(frames, batch):
    """Fewer rows than a 32-row T32 block / a 128-row tile at every stage: HiFi-GAN V1 and iSTFTNet, fp32
    path against the PyTorch ops, 16-bit path SNR."""
    from oracle import torch_port
    for cfg, fwd in ((synth.HIFIGAN_V1, torch_port.hifigan_forward), (synth.ISTFTNET, torch_port.istftnet_forward)):
        state = synth.make_state(cfg, 11, "unit")
        gen = build_generator(cfg, state, DEV, True)
        mel = torch.from_numpy(synth.make_mel(batch, frames, 50 + frames)).to(DEV)
        folded = {k: v.to(DEV) for k, v in torch_port.fold_state(state).items()}
        with torch.no_grad():
            ref = fwd(folded, cfg, mel).reshape(batch, -1)
            gen.precision = "fp32"
            o32 = gen(mel).reshape(batch, -1)
            gen.precision = "bf16"
            o16 = gen(mel).reshape(batch, -1)
        assert o32.shape == ref.shape == o16.shape
        assert float((o32 - ref).abs().max()) <= 1e-4
        assert np_oracle.snr_db(ref.cpu().numpy(), o16.cpu().numpy()) >= 40.0
    assert not lib_mod.tc_abort_status()


@pytest.mark.gpu
def test_reassigned_parameter_is_picked_up():
    """Replacing a Parameter OBJECT (not just its data) must reload the handle on the next call."""
    cfg = synth.CONFIGS["hifigan_small"]
    gen = build_generator(cfg, synth.make_state(cfg, 5, "unit"), "cuda", remove_wn=True)
    gen.precision = "fp32"
    mel = torch.from_numpy(synth.make_mel(1, 6, 3)).cuda()
    with torch.no_grad():
        y0 = gen(mel)
        gen.conv_post.bias = torch.nn.Parameter(gen.conv_post.bias.detach() + 0.25)   # same shape, new object
        y1 = gen(mel)
        gen.conv_post.bias -= 0.25                                                    # in-place under no_grad: version bump only
        y2 = gen(mel)
    assert not torch.equal(y0, y1) and float((y1 - y0).abs().max()) > 1e-2
    assert float((y2 - y0).abs().max()) < 1e-6 and float((y2 - y1).abs().max()) > 1e-2   # (b + 0.25) - 0.25 rounds


@pytest.mark.parametrize("frames,batch", [(37, 2), (129, 3)])
def test_persistent_upsampler_matches_polyphase_launches(frames, batch, tmp_path):
    """ups_tc.cu (all phases of a ConvTranspose1d side by side in one persistent launch, transposed stores) against the
    per-phase conv_tc launches it replaces (NVSE_UPS_TC=0, read once per process: run in a child): same half operands and
    fp32 accumulation, only the summation order differs.  Tile counts with partial last tiles at every stage."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "polyphase.npy")
    child = (
        "import sys, numpy as np, torch\n"
        f"sys.path[:0] = [{here!r}, {os.path.dirname(here)!r}]\n"
        "import synth\n"
        "from util import build_generator\n"
        "cfg = synth.HIFIGAN_V1\n"
        "gen = build_generator(cfg, synth.make_state(cfg, 21, 'init'), 'cuda:0', True)\n"
        f"mel = torch.from_numpy(synth.make_mel({batch}, {frames}, 77)).to('cuda:0')\n"
        "with torch.no_grad():\n"
        f"    np.save({out!r}, gen(mel).reshape({batch}, -1).cpu().numpy())\n"
    )
    env = dict(os.environ, NVSE_UPS_TC="0")
    subprocess.run([sys.executable, "-c", child], check=True, env=env, timeout=600)
    ref = np.load(out)
    gen = build_generator(A, synth.make_state(A, 21, "init"), DEV, True)
    mel = torch.from_numpy(synth.make_mel(batch, frames, 77)).to(DEV)
    with torch.no_grad():
        got = gen(mel).reshape(batch, -1).cpu().numpy()
        gen.precision = "fp32"
        f32 = gen(mel).reshape(batch, -1).cpu().numpy()
    assert got.shape == ref.shape
    # a different fp32 summation order moves some 16-bit roundings of the ResBlocks behind the upsamplers: the two paths
    # agree far better than either agrees with the fp32 path, and are equally close to it
    snr = np_oracle.snr_db(ref, got)
    s_new, s_old = np_oracle.snr_db(f32, got), np_oracle.snr_db(f32, ref)
    report(f"persistent upsampler vs polyphase launches ({batch} x {frames} frames): SNR {snr:.1f} dB; vs the fp32 path {s_new:.1f} dB "
           f"(polyphase launches: {s_old:.1f} dB)")
    assert snr >= 60.0 and snr >= max(s_new, s_old) + 3.0
    assert abs(s_new - s_old) <= 1.5 and s_new >= 40.0
    assert not lib_mod.tc_abort_status()
