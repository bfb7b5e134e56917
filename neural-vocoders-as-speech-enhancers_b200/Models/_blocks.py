"""Parameter containers shared by HiFiGAN and iSTFTNet.

The sub-module tree mirrors the reference's attribute names (``conv_pre``, ``ups``,
``resblocks.N.convs1.M`` ...) because those names ARE the checkpoint wire format
(``*.weight_g`` / ``*.weight_v`` / ``*.bias``; train_time_wi_inv.py:254,
infers/inference_hifigan.py:45).  None of these torch modules is ever called: the
arithmetic runs in the sm_100a kernels behind ``_engine.GeneratorEngine``."""
from __future__ import annotations

import warnings

import torch
from torch import nn

from .. import _lib
from .._engine import GeneratorEngine


def _wn(conv):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        return nn.utils.weight_norm(conv)  # old-style: weight_g / weight_v keys, like hifigan.py:5


def _strip_wn(conv):
    nn.utils.remove_weight_norm(conv)


def _redraw_like_reference(container):
    """The reference runs ``init_weights`` (N(0, 0.01) on ``m.weight.data``) over these
    sub-trees (hifigan.py:31,41,105-106).  Under old-style weight_norm ``m.weight`` is a
    derived tensor, so the call leaves weight_g / weight_v untouched but DOES advance the
    global RNG; it is replayed here so that ``torch.manual_seed(s); HiFiGAN(h)`` yields the
    same parameters as the reference constructor."""
    for m in container.modules():
        if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
            m.weight.data.normal_(0.0, 0.01)


def _same_pad(kernel_size, dilation):
    return (kernel_size - 1) * dilation // 2


class ResBlock1(nn.Module):
    """Three (dilated conv, conv) pairs with residuals -- hifigan.py:19-56."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = nn.ModuleList(
            _wn(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=_same_pad(kernel_size, d)))
            for d in (dilation[0], dilation[1], dilation[2]))
        _redraw_like_reference(self.convs1)
        self.convs2 = nn.ModuleList(
            _wn(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=_same_pad(kernel_size, 1)))
            for _ in range(3))
        _redraw_like_reference(self.convs2)

    def remove_weight_norm(self):
        for conv in list(self.convs1) + list(self.convs2):
            _strip_wn(conv)


class ResBlock2(nn.Module):
    """Two dilated convs with residuals -- hifigan.py:59-80."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = nn.ModuleList(
            _wn(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=_same_pad(kernel_size, d)))
            for d in (dilation[0], dilation[1]))
        _redraw_like_reference(self.convs)

    def remove_weight_norm(self):
        for conv in self.convs:
            _strip_wn(conv)


class _GeneratorBase(nn.Module):
    """Common constructor of the HiFi-GAN-family generators (hifigan.py:84-106,
    istftnet.py:272-297): conv_pre, transposed-conv upsamplers, MRF resblocks, conv_post."""

    _kind = None

    def __init__(self, h, post_channels):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        c0 = h.upsample_initial_channel
        self.conv_pre = _wn(nn.Conv1d(80, c0, 7, 1, padding=3))
        block = ResBlock1 if h.resblock == "1" else ResBlock2
        self.ups = nn.ModuleList(
            _wn(nn.ConvTranspose1d(c0 // 2 ** i, c0 // 2 ** (i + 1), k, u, padding=(k - u) // 2))
            for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // 2 ** (i + 1)
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(block(h, ch, k, d))
        self.conv_post = _wn(nn.Conv1d(ch, post_channels, 7, 1, padding=3))
        _redraw_like_reference(self.ups)
        _redraw_like_reference(self.conv_post)
        # inference: "bf16" (tcgen05 tensor cores, default) or "fp32"; also NVSE_B200_PRECISION in the environment
        self.precision = None
        # training: fp32 like the reference unless set to "bf16" (or NVSE_B200_TRAIN_PRECISION / an explicit .precision)
        self.train_precision = None
        self._engine = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_engine())

    def invalidate_engine(self):
        """The native handle re-reads every parameter at the next call (see GeneratorEngine.invalidate)."""
        if self._engine is not None:
            self._engine.invalidate()

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): parameters may be re-created
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_engine()
        return out

    def forward(self, x, out=None, frames=None):
        """mel ``[B, 80, frames]`` -> waveform ``[B, samples]`` on ``x.device`` (``out``: inference only, write into this
        device tensor instead of allocating the result; ``frames``: inference only, mel frames per utterance of a padded
        batch -- every utterance is computed exactly as if it were passed alone, see ``GeneratorEngine.forward``)."""
        if self._engine is None:
            object.__setattr__(self, "_engine", GeneratorEngine(self, self._kind))
        return self._engine.forward(self, x, out=out, frames=frames)

    @torch.no_grad()
    def forward_pcm16(self, x, out=None):
        """Inference straight to 16-bit PCM: what the reference's loop does with the waveform next (``sf.write(path,
        audio, sr, 'PCM_16')``, infers/inference_hifigan.py:89-95), with the quantisation fused into the last kernel.
        mel ``[B, 80, frames]`` -> int16 ``[B, samples]``."""
        if self._engine is None:
            object.__setattr__(self, "_engine", GeneratorEngine(self, self._kind))
        return self._engine.forward(self, x, pcm16=True, out=out)

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for up in self.ups:
            _strip_wn(up)
        for block in self.resblocks:
            block.remove_weight_norm()
        _strip_wn(self.conv_pre)
        _strip_wn(self.conv_post)
        self.invalidate_engine()

    def __getstate__(self):  # the native handle is process-local
        state = self.__dict__.copy()
        state["_engine"] = None
        return state


GEN_HIFIGAN, GEN_ISTFTNET = _lib.GEN_HIFIGAN, _lib.GEN_ISTFTNET
