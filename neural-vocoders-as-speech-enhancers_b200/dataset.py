"""Drop-in for the reference ``dataset.mel_spectrogram`` (dataset.py:53-91), computed by
the fused sm_100a front-end kernel (csrc/frontend.cu) through the C ABI.

Same signature, same positional order, same module-level caches (``mel_window``,
``param_string``) as the reference, so ``from dataset import mel_spectrogram`` call sites
(infers/inference_hifigan.py:31-32, train_time_wi_inv.py:177-186, Models/models.py:164)
work unchanged.  Results live on the device the reference would have used: ``y.device``
or CPU when ``in_dataset`` -- CPU inputs are staged through the GPU, there is no CPU
implementation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .melbasis import slaney_mel_basis

mel_window = {}      # param_string -> (mel_basis, hann_window) tensors, as in dataset.py:45
inv_mel_window = {}  # kept for dataset.inverse_mel compatibility (dataset.py:46)
_frontends = {}      # (param_string without device, hop, cuda index) -> nvse_frontend*
_istft_idents = []   # identity keys of window tensors that alias entries of _frontends (istft)


def param_string(sampling_rate, n_fft, num_mels, fmin, fmax, win_size, device):
    """dataset.py:49-50."""
    return f"{sampling_rate}-{n_fft}-{num_mels}-{fmin}-{fmax}-{win_size}-{device}"


def dynamic_range_compression_torch(x, C=1, clip_val=1e-5):
    """dataset.py:27-28 (kept for callers; the kernel applies it in its epilogue)."""
    return torch.log(torch.clamp(x, min=clip_val) * C)


def spectral_normalize_torch(magnitudes):
    """dataset.py:35-37."""
    return dynamic_range_compression_torch(magnitudes)


def _cuda_device_for(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _lib.NvseError("mel_spectrogram needs a CUDA device: the B200 front-end has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _frontend(sampling_rate, n_fft, num_mels, hop_size, win_size, fmin, fmax, basis, window, dev):
    key = (sampling_rate, n_fft, num_mels, fmin, fmax, win_size, hop_size, dev.index)
    fe = _frontends.get(key)
    if fe is None:
        lib = _lib.load()
        win = window.detach().to("cpu", torch.float32)
        if win_size < n_fft:  # torch.stft centres a short window inside n_fft
            left = (n_fft - win_size) // 2
            win = torch.nn.functional.pad(win, (left, n_fft - win_size - left))
        win = np.ascontiguousarray(win.numpy())
        mb = np.ascontiguousarray(basis.detach().to("cpu", torch.float32).numpy())
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.nvse_frontend_create(n_fft, hop_size, num_mels, win.ctypes.data_as(C.c_void_p),
                                                mb.ctypes.data_as(C.c_void_p), C.byref(handle)))
        fe = _frontends[key] = handle
    return fe


def frontend_handle(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, dev):
    """The native front-end handle (``nvse_frontend*``) of this parameter set on CUDA device ``dev`` -- what ``mel_spectrogram``
    uses and what the fused wav -> wav call (``nvse_vocoder_forward``, pipeline.Vocoder) takes."""
    ps = param_string(sampling_rate, n_fft, num_mels, fmin, fmax, win_size, dev)
    if ps in mel_window:
        mel_basis, hann_window = mel_window[ps]
    else:
        mel_basis = torch.from_numpy(slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)).float().to(dev)
        hann_window = torch.hann_window(win_size).to(dev)
        mel_window[ps] = (mel_basis, hann_window)
    return _frontend(sampling_rate, n_fft, num_mels, hop_size, win_size, fmin, fmax, mel_basis, hann_window, dev)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center=True, in_dataset=False, *, lengths=None):
    """log-mel spectrogram ``[B, num_mels, 1 + T // hop_size]`` (or ``[num_mels, F]`` for 1-D
    input) of ``y``.  ``center`` is accepted and ignored exactly as in the reference, which
    always calls ``torch.stft(center=True)`` (dataset.py:62 vs :84).

    ``lengths`` (extension, keyword-only, inference): int tensor ``[B]`` of samples per utterance for a batch padded to
    the longest one.  Utterance ``b`` gets the log-mel of its own ``lengths[b]`` samples (reflect padding at its own end),
    bit-identical to a single-utterance call, in frames ``[0, 1 + lengths[b] // hop_size)``; later frames are undefined."""
    global mel_window
    if y.dim() not in (1, 2):
        raise RuntimeError(f"mel_spectrogram expects a 1-D or 2-D waveform tensor, got {tuple(y.shape)}")
    if lengths is not None and (y.dim() != 2 or (torch.is_grad_enabled() and y.requires_grad)):
        raise RuntimeError("mel_spectrogram(lengths=...) takes a 2-D batch and is an inference feature")
    if torch.is_grad_enabled() and y.requires_grad:  # the mel-L1 loss of the trainer (train_time_wi_inv.py:173-179,231-235)
        return _MelFn.apply(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset)
    out_device = torch.device("cpu") if in_dataset else y.device
    ps = param_string(sampling_rate, n_fft, num_mels, fmin, fmax, win_size, out_device)
    if ps in mel_window:
        mel_basis, hann_window = mel_window[ps]
    else:
        mel_basis = torch.from_numpy(slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)).float().to(out_device)
        hann_window = torch.hann_window(win_size).to(out_device)
        mel_window[ps] = (mel_basis, hann_window)

    dev = _cuda_device_for(y)
    fe = _frontend(sampling_rate, n_fft, num_mels, hop_size, win_size, fmin, fmax, mel_basis, hann_window, dev)
    squeeze = y.dim() == 1
    yd = y.detach().to(dev, torch.float32)
    if squeeze:
        yd = yd.unsqueeze(0)
    if yd.stride(-1) != 1:
        yd = yd.contiguous()
    batch, samples = yd.shape
    frames = 1 + samples // hop_size
    out = torch.empty((batch, num_mels, frames), dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        if lengths is not None:
            ld = torch.as_tensor(lengths).to(dev, torch.int32).contiguous()
            if ld.shape != (batch,):
                raise RuntimeError(f"lengths must have shape [{batch}], got {tuple(ld.shape)}")
            if batch and (int(ld.min()) <= n_fft // 2 or int(ld.max()) > samples):
                raise RuntimeError(f"lengths must lie in ({n_fft // 2}, {samples}] (reflect padding needs more than n_fft/2 samples)")
            out.zero_()  # frames beyond an utterance's own are not written by the kernel
            _lib.check(lib.nvse_frontend_mel_ragged_f32(fe, _lib.ptr(yd), batch, samples, yd.stride(0) if batch > 1 else samples,
                                                        _lib.ptr(ld), _lib.ptr(out), C.c_void_p(stream)))
        else:
            _lib.check(lib.nvse_frontend_mel_f32(fe, _lib.ptr(yd), batch, samples, yd.stride(0) if batch > 1 else samples,
                                                 _lib.ptr(out), C.c_void_p(stream)))
    if squeeze:
        out = out[0]
    return out if out.device == out_device else out.to(out_device)


def spectral_de_normalize_torch(magnitudes):
    """dataset.py:39-41 (kept for callers)."""
    return torch.exp(magnitudes)


def inverse_mel(mel, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset=False):
    """Reference ``dataset.inverse_mel`` (dataset.py:94-121): ``pinv(mel_basis) @ exp(mel)`` -> ``[B, n_fft//2+1, F]``
    (or ``[n_fft//2+1, F]`` for 2-D input), with the reference's ``inv_mel_window`` / ``mel_window`` caches."""
    global inv_mel_window, mel_window
    if mel.dim() not in (2, 3):
        raise RuntimeError(f"inverse_mel expects [B, num_mels, F] or [num_mels, F], got {tuple(mel.shape)}")
    out_device = torch.device("cpu") if in_dataset else mel.device
    ps = param_string(sampling_rate, n_fft, num_mels, fmin, fmax, win_size, out_device)
    if ps in inv_mel_window:
        inv_basis = inv_mel_window[ps]
    else:
        if ps in mel_window:
            mel_basis, _ = mel_window[ps]
        else:
            mel_basis = torch.from_numpy(slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)).float().to(out_device)
            mel_window[ps] = (mel_basis, torch.hann_window(win_size).to(out_device))
        inv_basis = mel_basis.pinverse()          # once per parameter set, like dataset.py:118
        inv_mel_window[ps] = inv_basis
    dev = _cuda_device_for(mel)
    squeeze = mel.dim() == 2
    md = mel.detach().to(dev, torch.float32)
    if squeeze:
        md = md.unsqueeze(0)
    md = md.contiguous()
    batch, n_mels, frames = md.shape
    if n_mels != inv_basis.shape[1]:
        raise RuntimeError(f"inverse_mel: mel has {n_mels} bands, the basis {inv_basis.shape[1]}")
    ib = inv_basis.to(dev, torch.float32).contiguous()
    out = torch.empty((batch, ib.shape[0], frames), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.load().nvse_inverse_mel_f32(_lib.ptr(ib), _lib.ptr(md), _lib.ptr(out), batch, ib.shape[0], n_mels,
                                                    frames, stream))
    if squeeze:
        out = out[0]
    return out if out.device == out_device else out.to(out_device)


def amp_pha_specturm(y, n_fft, hop_size, win_size):
    """Reference ``dataset.amp_pha_specturm`` (dataset.py:124-139; the spelling is the reference's):
    ``(log(|X| + 1e-7), atan2(Im X, Re X), Re X, Im X)`` of ``torch.stft(y, center=True)``, each ``[B, n_fft//2+1, F]``."""
    if y.dim() not in (1, 2):
        raise RuntimeError(f"amp_pha_specturm expects a 1-D or 2-D waveform tensor, got {tuple(y.shape)}")
    dev = _cuda_device_for(y)
    fe = _frontends.get(("stft", n_fft, 1, 0, 0, win_size, hop_size, dev.index))
    if fe is None:  # a front-end handle without a mel basis (one zero filter): window + twiddles only
        fe = _frontend("stft", n_fft, 1, hop_size, win_size, 0, 0, torch.zeros(1, n_fft // 2 + 1), torch.hann_window(win_size), dev)
    squeeze = y.dim() == 1
    yd = y.detach().to(dev, torch.float32)
    if squeeze:
        yd = yd.unsqueeze(0)
    if yd.stride(-1) != 1:
        yd = yd.contiguous()
    batch, samples = yd.shape
    frames = 1 + samples // hop_size
    outs = [torch.empty((batch, n_fft // 2 + 1, frames), dtype=torch.float32, device=dev) for _ in range(4)]
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.load().nvse_frontend_stft_f32(fe, _lib.ptr(yd), batch, samples, yd.stride(0) if batch > 1 else samples,
                                                      *[_lib.ptr(o) for o in outs], stream))
    if squeeze:
        outs = [o[0] for o in outs]
    return tuple(o if o.device == y.device else o.to(y.device) for o in outs)


def istft(spec, n_fft, hop_length=None, win_length=None, window=None, center=True):
    """The inverse STFT the reference's T-F vocoders end in (Models/apnet.py:155, Models/freeV.py:178, Models/bsrnn.py:210):
    ``torch.istft(spec, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=True)`` for a
    complex ``spec`` of shape ``[B, n_fft//2+1, F]`` (or ``[n_fft//2+1, F]``) -> ``[B, hop * (F - 1)]`` float32.
    Same argument order as ``torch.istft``, so the call sites change one name.  n_fft = 1024 (this build's FFT size),
    ``center=True`` only; ``window`` defaults to the periodic Hann window those call sites pass."""
    if not center:
        raise NotImplementedError("istft: the reference only calls torch.istft with center=True")
    if not torch.is_complex(spec) or spec.dim() not in (2, 3):
        raise RuntimeError(f"istft expects a complex [B, n_fft//2+1, F] or [n_fft//2+1, F] tensor, got {spec.dtype} {tuple(spec.shape)}")
    hop_length = n_fft // 4 if hop_length is None else int(hop_length)
    win_length = n_fft if win_length is None else int(win_length)
    if spec.shape[-2] != n_fft // 2 + 1:
        raise RuntimeError(f"istft: expected {n_fft // 2 + 1} frequency bins, got {spec.shape[-2]}")
    dev = _cuda_device_for(spec)
    # the handle of a window tensor seen before is found by its identity (storage address + version): hashing its values means
    # a device -> host copy and a synchronisation on every call, which is what the reference's call sites would pay
    ident = None if window is None else ("istft-window", window.data_ptr(), window._version, tuple(window.shape), str(window.device),
                                         n_fft, win_length, hop_length, dev.index)
    fe = _frontends.get(ident) if ident is not None else None
    if fe is None:
        win = torch.hann_window(win_length) if window is None else window.detach().to("cpu", torch.float32)
        key = ("istft", n_fft, 1, hash(win.numpy().tobytes()), 0, win_length, hop_length, dev.index)
        fe = _frontends.get(key)
        if fe is None:  # a front-end handle without a mel basis: window + twiddles only
            _frontend("istft", n_fft, 1, hop_length, win_length, hash(win.numpy().tobytes()), 0, torch.zeros(1, n_fft // 2 + 1), win, dev)
            fe = _frontends[key]
        if ident is not None:
            if len(_istft_idents) > 64:  # bounded: windows that are re-created on every call must not grow the table
                for k in _istft_idents:
                    _frontends.pop(k, None)
                _istft_idents.clear()
            _frontends[ident] = fe
            _istft_idents.append(ident)
    squeeze = spec.dim() == 2
    sd = spec.detach().to(dev)
    if squeeze:
        sd = sd.unsqueeze(0)
    if sd.dtype != torch.complex64:
        sd = sd.to(torch.complex64)
    sd = sd.contiguous()             # torch's complex64 layout (interleaved re, im) is read by the kernel as it is
    cview = torch.view_as_real(sd)   # [B, bins, F, 2] float32, same storage
    batch, _, frames = sd.shape
    out = torch.empty((batch, hop_length * (frames - 1)), dtype=torch.float32, device=dev)
    lib = _lib.load()
    scratch = torch.empty(lib.nvse_frontend_istft_scratch_bytes(fe, batch, frames), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.nvse_frontend_istft_c64(fe, _lib.ptr(cview), batch, frames, _lib.ptr(out), _lib.ptr(scratch),
                                               scratch.numel(), stream))
    if squeeze:
        out = out[0]
    return out if out.device == spec.device else out.to(spec.device)


class _MelFn(torch.autograd.Function):
    """mel_spectrogram as an autograd node: the CUDA forward above, and the CUDA backward
    (nvse_frontend_mel_backward_f32: recomputed spectra -> d|X| -> packed inverse FFT -> overlap-add)."""

    @staticmethod
    def forward(ctx, y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset):
        with torch.no_grad():
            out = mel_spectrogram(y.detach(), n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                                  in_dataset=in_dataset)
        ctx.args = (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset)
        ctx.save_for_backward(y)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dmel):
        (y,) = ctx.saved_tensors
        n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset = ctx.args
        out_device = torch.device("cpu") if in_dataset else y.device
        mel_basis, hann_window = mel_window[param_string(sampling_rate, n_fft, num_mels, fmin, fmax, win_size, out_device)]
        dev = _cuda_device_for(y)
        fe = _frontend(sampling_rate, n_fft, num_mels, hop_size, win_size, fmin, fmax, mel_basis, hann_window, dev)
        squeeze = y.dim() == 1
        yd = y.detach().to(dev, torch.float32)
        gd = dmel.detach().to(dev, torch.float32)
        if squeeze:
            yd, gd = yd.unsqueeze(0), gd.unsqueeze(0)
        yd, gd = yd.contiguous(), gd.contiguous()
        batch, samples = yd.shape
        dy = torch.empty_like(yd)
        lib = _lib.load()
        scratch = torch.empty(lib.nvse_frontend_backward_scratch_bytes(fe, batch, samples), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.nvse_frontend_mel_backward_f32(fe, _lib.ptr(yd), batch, samples, samples, _lib.ptr(gd), _lib.ptr(dy),
                                                          _lib.ptr(scratch), scratch.numel(), C.c_void_p(stream)))
        if squeeze:
            dy = dy[0]
        return (dy.to(y.device, y.dtype),) + (None,) * 8
