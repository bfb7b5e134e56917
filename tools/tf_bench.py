#!/usr/bin/env python
"""amp_pha_specturm / inverse_mel at the cfg2 shape (64 x 4 s): CUDA-event timed, algorithmic GB/s, and stock PyTorch
(the reference's own expression, dataset.py:94-139) on the same GPU."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg  # noqa: E402
import synth  # noqa: E402
a = synth.HIFIGAN_V1
B, T = 64, 88200
y = torch.from_numpy(synth.make_wave(B, T, 0)).cuda()
F = 1 + T // 256


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


win = torch.hann_window(1024, device="cuda")
def stock_stft():
    s = torch.stft(y, 1024, hop_length=256, win_length=1024, window=win, center=True, return_complex=True)
    return torch.log(s.abs() + 1e-7), torch.atan2(s.imag, s.real), s.real, s.imag
def kernel_ms(fn, name):
    pkg._lib.profile_begin()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    return sum(k["ms"] for k in pkg._lib.profile_end() if k["kernel"].startswith(name)) / 10


ms = timed(lambda: pkg.amp_pha_specturm(y, 1024, 256, 1024))
ms_t = timed(stock_stft)
print(f"  kernel alone: {kernel_ms(lambda: pkg.amp_pha_specturm(y, 1024, 256, 1024), 'stft_amp_pha'):.3f} ms")
byt = 4 * B * T + 4 * 4 * B * 513 * F
print(f"amp_pha_specturm {B} x {T}: {ms:.3f} ms ({byt / ms / 1e6:.0f} GB/s algorithmic: waveform in + 4 planes out) vs stock torch {ms_t:.3f} ms -> {ms_t / ms:.2f}x")
margs = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["fmax"])
mel = pkg.mel_spectrogram(y, *margs)
basis = torch.from_numpy(pkg.melbasis.slaney_mel_basis(a["sampling_rate"], a["n_fft"], a["num_mels"], a["fmin"], a["fmax"])).cuda()
inv = basis.pinverse()
print(f"  kernel alone: {kernel_ms(lambda: pkg.inverse_mel(mel, *margs), 'inverse_mel'):.3f} ms")
ms = timed(lambda: pkg.inverse_mel(mel, *margs))
ms_t = timed(lambda: inv @ torch.exp(mel))
byt = 4 * B * F * (80 + 513)
print(f"inverse_mel {B} x 80 x {F}: {ms:.3f} ms ({byt / ms / 1e6:.0f} GB/s algorithmic) vs stock torch {ms_t:.3f} ms -> {ms_t / ms:.2f}x")

# n_fft = 1024 inverse STFT (the T-F vocoders' head) against torch.istft on the same GPU
spec = torch.stft(y, 1024, hop_length=256, win_length=1024, window=win, center=True, return_complex=True)
ms = timed(lambda: pkg.istft(spec, 1024, 256, 1024, win))
ms_t = timed(lambda: torch.istft(spec, 1024, hop_length=256, win_length=1024, window=win, center=True))
err = float((pkg.istft(spec, 1024, 256, 1024, win) - torch.istft(spec, 1024, hop_length=256, win_length=1024, window=win, center=True)).abs().max())
print(f"  kernels alone: {kernel_ms(lambda: pkg.istft(spec, 1024, 256, 1024, win), 'istft1024'):.3f} ms")
print(f"istft {B} x 513 x {F}: {ms:.3f} ms vs stock torch.istft {ms_t:.3f} ms -> {ms_t / ms:.2f}x  (max |difference| {err:.1e})")
