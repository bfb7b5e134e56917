"""float64 numpy restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines (relative to /root/reference) it restates.
Nothing here is imported by the product package; see ``oracle/__init__.py``.

Layouts follow the reference (PyTorch) conventions: activations ``[B, C, T]``,
Conv1d weights ``[Cout, Cin, k]``, ConvTranspose1d weights ``[Cin, Cout, k]``.
"""
from __future__ import annotations

import numpy as np

LRELU_SLOPE = 0.1  # Models/hifigan.py:7


# ----------------------------------------------------------------------------------
# mel filterbank -- librosa.filters.mel (librosa==0.10.2.post1, requirements.txt:4),
# called at dataset.py:73.  Third-party, not vendored: published algorithm restated.
# ----------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_t, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with library defaults
    (htk=False, norm='slaney', dtype=float32).  Returns float32 ``[n_mels, n_fft//2+1]``.

    The float32 rounding points of librosa are reproduced: the triangular weight is
    stored into a float32 array, then scaled in place by the float64 Slaney factor.
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights = (weights.astype(np.float64) * enorm[:, None]).astype(np.float32)
    return weights


# ----------------------------------------------------------------------------------
# mel_spectrogram -- dataset.py:53-91
# ----------------------------------------------------------------------------------
def hann_periodic(n):
    """torch.hann_window(n) (periodic=True default), dataset.py:75."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def reflect_pad(y, pad):
    """torch.stft(center=True, pad_mode='reflect') padding, dataset.py:78-86."""
    y = np.asarray(y)
    if y.shape[-1] <= pad:
        raise ValueError("reflect padding needs T > n_fft // 2")
    left = y[..., 1 : pad + 1][..., ::-1]
    right = y[..., -pad - 1 : -1][..., ::-1]
    return np.concatenate([left, y, right], axis=-1)


def stft_magnitude(y, n_fft, hop, win, window=None):
    """|torch.stft(y, n_fft, hop, win, window, center=True)| -> ``[B, n_fft//2+1, F]``,
    F = 1 + T // hop.  dataset.py:78-88."""
    y = np.asarray(y, dtype=np.float64)
    squeeze = y.ndim == 1
    if squeeze:
        y = y[None]
    if window is None:
        window = hann_periodic(win)
    window = np.asarray(window, dtype=np.float64)
    if win < n_fft:  # torch.stft centres a short window inside n_fft
        lp = (n_fft - win) // 2
        window = np.pad(window, (lp, n_fft - win - lp))
    yp = reflect_pad(y, n_fft // 2)
    n_frames = 1 + y.shape[-1] // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[:, idx] * window  # [B, F, n_fft]
    spec = np.fft.rfft(frames, axis=-1)  # unnormalised, onesided
    mag = np.abs(spec).transpose(0, 2, 1)
    return mag[0] if squeeze else mag


def stft_complex(y, n_fft, hop, win):
    """torch.stft(y, n_fft, hop, win, hann, center=True, return_complex=True) -> complex128 ``[B, n_fft//2+1, F]``."""
    y = np.asarray(y, dtype=np.float64)
    squeeze = y.ndim == 1
    if squeeze:
        y = y[None]
    window = hann_periodic(win)
    if win < n_fft:
        lp = (n_fft - win) // 2
        window = np.pad(window, (lp, n_fft - win - lp))
    yp = reflect_pad(y, n_fft // 2)
    n_frames = 1 + y.shape[-1] // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(yp[:, idx] * window, axis=-1).transpose(0, 2, 1)
    return spec[0] if squeeze else spec


def amp_pha_spectrum(y, n_fft, hop, win):
    """dataset.py:124-139 (``amp_pha_specturm``): (log(|X| + 1e-7), atan2(Im, Re), Re, Im)."""
    spec = stft_complex(y, n_fft, hop, win)
    return np.log(np.abs(spec) + 1e-7), np.arctan2(spec.imag, spec.real), spec.real, spec.imag


def istft(spec, n_fft, hop, win):
    """torch.istft(spec, n_fft, hop, win, window=hann_window(win), center=True) -- the head of the reference's T-F vocoders
    (Models/apnet.py:155, Models/freeV.py:178, Models/bsrnn.py:210): per-frame irfft (1/n_fft), times the window,
    overlap-add, divided by the overlap-added squared window, n_fft//2 samples trimmed on both sides -> ``[B, hop*(F-1)]``."""
    spec = np.asarray(spec, dtype=np.complex128)
    squeeze = spec.ndim == 2
    if squeeze:
        spec = spec[None]
    window = hann_periodic(win)
    if win < n_fft:
        lp = (n_fft - win) // 2
        window = np.pad(window, (lp, n_fft - win - lp))
    B, _, F = spec.shape
    frames = np.fft.irfft(spec.transpose(0, 2, 1), n=n_fft, axis=-1) * window  # [B, F, n_fft]
    total = n_fft + hop * (F - 1)
    y = np.zeros((B, total))
    env = np.zeros(total)
    for f in range(F):
        y[:, f * hop:f * hop + n_fft] += frames[:, f]
        env[f * hop:f * hop + n_fft] += window ** 2
    lo, hi = n_fft // 2, n_fft // 2 + hop * (F - 1)
    out = y[:, lo:hi] / env[lo:hi]
    return out[0] if squeeze else out


def inverse_mel(mel, inv_basis):
    """dataset.py:94-121: ``inv_basis @ exp(mel)`` with ``inv_basis = pinverse(mel_basis)`` supplied by the caller
    (the reference takes it from ``torch.Tensor.pinverse``; the checker does not re-derive a pseudo-inverse)."""
    return np.matmul(np.asarray(inv_basis, dtype=np.float64), np.exp(np.asarray(mel, dtype=np.float64)))


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center=True, mel_basis=None):
    """dataset.py:53-91: log(clamp(mel_basis @ |STFT|, 1e-5)).  float32 result.
    ``center`` is accepted and ignored, as in the reference (dataset.py:62 vs :84)."""
    if mel_basis is None:
        mel_basis = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
    mag = stft_magnitude(y, n_fft, hop_size, win_size)
    mel = np.matmul(np.asarray(mel_basis, dtype=np.float64), mag)
    return np.log(np.maximum(mel, 1e-5)).astype(np.float32)  # dataset.py:27-28


# ----------------------------------------------------------------------------------
# layers -- Models/hifigan.py:9-80
# ----------------------------------------------------------------------------------
def get_padding(kernel_size, dilation=1):
    """Models/hifigan.py:15-16."""
    return int((kernel_size * dilation - dilation) / 2)


def weight_norm_fold(v, g):
    """torch.nn.utils.weight_norm (dim=0) fold, used by remove_weight_norm at
    Models/hifigan.py:52-56,126-133: w = g * v / ||v||, norm over all dims but 0."""
    v = np.asarray(v, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    norm = np.sqrt((v * v).reshape(v.shape[0], -1).sum(axis=1)).reshape(g.shape)
    return v * (g / norm)


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * slope)


def conv1d(x, w, b, dilation=1, padding=0):
    """torch.nn.Conv1d(stride=1) forward. x [B,Cin,T], w [Cout,Cin,k], b [Cout]."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    bsz, cin, t = x.shape
    cout, cin_w, k = w.shape
    assert cin == cin_w
    t_out = t + 2 * padding - dilation * (k - 1)
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    out = np.zeros((bsz, cout, t_out), dtype=np.float64)
    for j in range(k):
        out += np.matmul(w[None, :, :, j], xp[:, :, j * dilation : j * dilation + t_out])
    if b is not None:
        out += np.asarray(b, dtype=np.float64)[None, :, None]
    return out


def conv_transpose1d(x, w, b, stride, padding):
    """torch.nn.ConvTranspose1d forward. x [B,Cin,T], w [Cin,Cout,k].
    y[co, stride*t + j - padding] += x[ci,t] * w[ci,co,j]  (Models/hifigan.py:93-96)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    bsz, cin, t = x.shape
    cin_w, cout, k = w.shape
    assert cin == cin_w
    full = np.zeros((bsz, cout, (t - 1) * stride + k), dtype=np.float64)
    for j in range(k):
        contrib = np.matmul(w[:, :, j].T[None], x)  # [B, Cout, T]
        full[:, :, j : j + (t - 1) * stride + 1 : stride] += contrib
    t_out = (t - 1) * stride - 2 * padding + k
    out = full[:, :, padding : padding + t_out]
    if b is not None:
        out = out + np.asarray(b, dtype=np.float64)[None, :, None]
    return out


def _folded(state, prefix):
    """Weight of a (possibly weight-normed) layer from a reference-format state dict:
    either ``prefix.weight`` or ``prefix.weight_g`` / ``prefix.weight_v``."""
    if prefix + ".weight" in state:
        w = np.asarray(state[prefix + ".weight"], dtype=np.float64)
    else:
        w = weight_norm_fold(state[prefix + ".weight_v"], state[prefix + ".weight_g"])
    return w, np.asarray(state[prefix + ".bias"], dtype=np.float64)


def resblock1(state, prefix, x, kernel_size, dilations):
    """Models/hifigan.py:43-50."""
    for m, d in enumerate(dilations):
        w1, b1 = _folded(state, f"{prefix}.convs1.{m}")
        w2, b2 = _folded(state, f"{prefix}.convs2.{m}")
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, w1, b1, dilation=d, padding=get_padding(kernel_size, d))
        xt = leaky_relu(xt, LRELU_SLOPE)
        xt = conv1d(xt, w2, b2, dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def resblock2(state, prefix, x, kernel_size, dilations):
    """Models/hifigan.py:71-76."""
    for m, d in enumerate(dilations):
        w, b = _folded(state, f"{prefix}.convs.{m}")
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, w, b, dilation=d, padding=get_padding(kernel_size, d))
        x = xt + x
    return x


def _trunk(state, cfg, mel):
    """conv_pre + upsample/MRF stages shared by HiFiGAN.forward (hifigan.py:109-119)
    and iSTFTNet.forward (istftnet.py:300-310)."""
    w, b = _folded(state, "conv_pre")
    x = conv1d(mel, w, b, padding=3)
    n_k = len(cfg["resblock_kernel_sizes"])
    block = resblock1 if str(cfg["resblock"]) == "1" else resblock2
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = leaky_relu(x, LRELU_SLOPE)
        w, b = _folded(state, f"ups.{i}")
        x = conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = block(state, f"resblocks.{i * n_k + j}", x, rk, rd)
            xs = r if xs is None else xs + r
        x = xs / n_k
    return x


def hifigan_forward(state, cfg, mel):
    """Models/hifigan.py:108-124.  mel [B,80,F] -> wav [B, prod(upsample_rates)*F]."""
    mel = np.asarray(mel, dtype=np.float64)
    x = _trunk(state, cfg, mel)
    x = leaky_relu(x, 0.01)  # F.leaky_relu default slope, hifigan.py:120
    w, b = _folded(state, "conv_post")
    x = conv1d(x, w, b, padding=3)
    return np.tanh(x)[:, 0, :]


def istft_head(mag, phase, n_fft, hop):
    """TorchSTFT.inverse, Models/istftnet.py:183-188: torch.istft(mag*exp(i*phase),
    n_fft, hop, n_fft, periodic-Hann window, center=True) -> [B, hop*(T'-1)]."""
    mag = np.asarray(mag, dtype=np.float64)
    phase = np.asarray(phase, dtype=np.float64)
    bsz, nb, tp = mag.shape
    assert nb == n_fft // 2 + 1
    spec = mag * np.exp(1j * phase)
    frames = np.fft.irfft(spec, n=n_fft, axis=1)  # [B, n_fft, T'], 1/n scaling
    win = hann_periodic(n_fft)  # scipy get_window('hann', fftbins=True), istftnet.py:173
    frames = frames * win[None, :, None]
    total = n_fft + hop * (tp - 1)
    out = np.zeros((bsz, total), dtype=np.float64)
    env = np.zeros(total, dtype=np.float64)
    for tau in range(tp):
        out[:, tau * hop : tau * hop + n_fft] += frames[:, :, tau]
        env[tau * hop : tau * hop + n_fft] += win * win
    start = n_fft // 2
    length = hop * (tp - 1)
    return out[:, start : start + length] / env[None, start : start + length]


def istftnet_forward(state, cfg, mel):
    """Models/istftnet.py:299-318."""
    mel = np.asarray(mel, dtype=np.float64)
    x = _trunk(state, cfg, mel)
    x = leaky_relu(x, 0.01)  # istftnet.py:311
    x = np.concatenate([x[:, :, 1:2], x], axis=2)  # ReflectionPad1d((1, 0)), istftnet.py:312
    w, b = _folded(state, "conv_post")
    x = conv1d(x, w, b, padding=3)
    n_fft = int(cfg["gen_istft_n_fft"])
    nb = n_fft // 2 + 1
    spec = np.exp(x[:, :nb, :])  # istftnet.py:314
    phase = np.sin(x[:, nb:, :])  # istftnet.py:315
    return istft_head(spec, phase, n_fft, int(cfg["gen_istft_hop_size"]))


# ----------------------------------------------------------------------------------
# metrics used by the parity gates
# ----------------------------------------------------------------------------------
def snr_db(ref, deg, demean=True):
    """Metrics/snr.py:25-31 (de-meaned SNR); ``demean=False`` gives the raw SNR."""
    ref = np.asarray(ref, dtype=np.float64).ravel()
    deg = np.asarray(deg, dtype=np.float64).ravel()
    if demean:
        ref = ref - ref.mean()
        deg = deg - deg.mean()
    return float(10 * np.log10(np.sum(ref ** 2) / np.sum((ref - deg) ** 2) + 1e-10))
