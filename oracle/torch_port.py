"""fp32 functional-PyTorch port of the reference hot path.  TEST INFRASTRUCTURE ONLY.

The reference is a PyTorch program whose CPU path is ``torch.stft`` + oneDNN
convolutions; this port issues the same library calls on the same shapes, from a
flat state dict instead of an ``nn.Module`` tree.  ``bench.py`` times it as the CPU
baseline (``cpu_baseline.kind == "port"``) and ``tests/`` use it as a second checker
next to ``np_oracle``.  Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import np_oracle

_basis_cache = {}


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center=True):
    """dataset.py:53-91 on CPU tensors."""
    key = (sampling_rate, n_fft, num_mels, fmin, fmax, win_size)
    if key not in _basis_cache:
        basis = torch.from_numpy(np_oracle.mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
        _basis_cache[key] = (basis, torch.hann_window(win_size))
    basis, window = _basis_cache[key]
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window,
                      center=True, return_complex=True)
    return torch.log(torch.clamp(basis @ spec.abs(), min=1e-5))


def fold_state(state):
    """remove_weight_norm (Models/hifigan.py:126-133) on a flat state dict:
    ``*.weight_g`` / ``*.weight_v`` pairs become ``*.weight``."""
    out = {}
    for name, t in state.items():
        t = torch.as_tensor(t, dtype=torch.float32)
        if name.endswith(".weight_v"):
            g = torch.as_tensor(state[name[:-1] + "g"], dtype=torch.float32)
            norm = t.reshape(t.shape[0], -1).norm(dim=1).reshape(g.shape)
            out[name[: -len("_v")]] = t * (g / norm)
        elif name.endswith(".weight_g"):
            continue
        else:
            out[name] = t
    return out


def _resblock(w, prefix, x, k, dilations, two_convs):
    for m, d in enumerate(dilations):
        if two_convs:  # ResBlock1, hifigan.py:43-50
            xt = F.conv1d(F.leaky_relu(x, 0.1), w[f"{prefix}.convs1.{m}.weight"],
                          w[f"{prefix}.convs1.{m}.bias"], dilation=d,
                          padding=np_oracle.get_padding(k, d))
            xt = F.conv1d(F.leaky_relu(xt, 0.1), w[f"{prefix}.convs2.{m}.weight"],
                          w[f"{prefix}.convs2.{m}.bias"], padding=np_oracle.get_padding(k, 1))
        else:  # ResBlock2, hifigan.py:71-76
            xt = F.conv1d(F.leaky_relu(x, 0.1), w[f"{prefix}.convs.{m}.weight"],
                          w[f"{prefix}.convs.{m}.bias"], dilation=d,
                          padding=np_oracle.get_padding(k, d))
        x = xt + x
    return x


def _trunk(w, cfg, mel):
    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    n_k = len(cfg["resblock_kernel_sizes"])
    two = str(cfg["resblock"]) == "1"
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(F.leaky_relu(x, 0.1), w[f"ups.{i}.weight"], w[f"ups.{i}.bias"],
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = _resblock(w, f"resblocks.{i * n_k + j}", x, rk, rd, two)
            xs = r if xs is None else xs + r
        x = xs / n_k
    return x


def hifigan_forward_autograd(folded, cfg, mel):
    """Models/hifigan.py:108-124, differentiable: the checker of the backward kernels (what
    ``loss.backward()`` computes through the generator in train_time_wi_inv.py:222-236)."""
    x = _trunk(folded, cfg, mel)
    x = F.conv1d(F.leaky_relu(x), folded["conv_post.weight"], folded["conv_post.bias"], padding=3)
    return torch.tanh(x).squeeze(1)


@torch.no_grad()
def hifigan_forward(folded, cfg, mel):
    """Models/hifigan.py:108-124 from a folded state dict (see ``fold_state``)."""
    return hifigan_forward_autograd(folded, cfg, mel)


def hifigan_gradients(state, cfg, mel, dout):
    """Gradients of ``(G(mel) * dout).sum()`` (G = HiFiGAN or iSTFTNet, by ``cfg["model_name"]``) w.r.t. every tensor of a reference-format state dict
    (weight_g / weight_v / bias, or folded weight / bias) and w.r.t. ``mel``, by torch autograd over the port.
    -> (out, {name: grad}, dmel)"""
    leaves = {k: torch.as_tensor(v, dtype=torch.float32).clone().requires_grad_(True) for k, v in state.items()}
    mel = torch.as_tensor(mel, dtype=torch.float32).clone().requires_grad_(True)
    fwd = istftnet_forward_autograd if cfg.get("model_name") == "iSTFTNet" else hifigan_forward_autograd
    out = fwd(fold_state(leaves), cfg, mel)
    (out * torch.as_tensor(dout, dtype=torch.float32)).sum().backward()
    return out.detach(), {k: v.grad for k, v in leaves.items()}, mel.grad


@torch.no_grad()
def istftnet_forward(folded, cfg, mel):
    """Models/istftnet.py:299-318."""
    return istftnet_forward_autograd(folded, cfg, mel)


def istftnet_forward_autograd(folded, cfg, mel):
    """Models/istftnet.py:299-318, differentiable (checker of the iSTFTNet backward kernels)."""
    x = _trunk(folded, cfg, mel)
    x = F.pad(F.leaky_relu(x), (1, 0), mode="reflect")
    x = F.conv1d(x, folded["conv_post.weight"], folded["conv_post.bias"], padding=3)
    n_fft = int(cfg["gen_istft_n_fft"])
    hop = int(cfg["gen_istft_hop_size"])
    nb = n_fft // 2 + 1
    spec = torch.exp(x[:, :nb]) * torch.exp(1j * torch.sin(x[:, nb:]))
    return torch.istft(spec, n_fft, hop, n_fft, window=torch.hann_window(n_fft, device=spec.device))


# ---- waveform discriminators (Models/models.py:15-113, 187-246): functional restatement over a flat state dict -------
def _folded_disc_weight(state, prefix):
    """The effective convolution weight of one layer: weight_norm (``weight_g * weight_v / ||weight_v||``, norm over all
    dims but 0 -- torch.nn.utils.weight_norm, Models/models.py:5,18) or the plain ``weight`` a test passes after folding."""
    if prefix + ".weight_g" in state:
        g, v = state[prefix + ".weight_g"], state[prefix + ".weight_v"]
        return g * v / v.flatten(1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
    return state[prefix + ".weight"]


DISC_P_LAYERS = [(3, 2), (3, 2), (3, 2), (3, 2), (1, 2)]             # (stride, padding) of DiscriminatorP.convs, k = 5
DISC_S_LAYERS = [(1, 1, 7), (2, 4, 20), (2, 16, 20), (4, 16, 20), (4, 16, 20), (1, 16, 20), (1, 1, 2)]  # (stride, groups, padding)


def disc_p_forward(state, prefix, x, period):
    """DiscriminatorP.forward (Models/models.py:63-87): reflect-pad to a multiple of the period, view as
    [B, 1, T/period, period], five (k,1) convolutions + leaky_relu(0.1), conv_post.  Returns (flattened logits, fmap)."""
    fmap = []
    if x.dim() == 2:
        x = x.unsqueeze(1)
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for i, (stride, pad) in enumerate(DISC_P_LAYERS):
        x = F.conv2d(x, _folded_disc_weight(state, f"{prefix}.convs.{i}"), state[f"{prefix}.convs.{i}.bias"],
                     stride=(stride, 1), padding=(pad, 0))
        x = F.leaky_relu(x, 0.1)
        fmap.append(x)
    x = F.conv2d(x, _folded_disc_weight(state, f"{prefix}.conv_post"), state[f"{prefix}.conv_post.bias"], padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def disc_s_forward(state, prefix, x):
    """DiscriminatorS.forward (Models/models.py:202-214) from folded / weight-normed weights (the spectral_norm variant is
    checked through the module itself, its weight depends on the power-iteration buffers)."""
    fmap = []
    if x.dim() == 2:
        x = x.unsqueeze(1)
    for i, (stride, groups, pad) in enumerate(DISC_S_LAYERS):
        x = F.conv1d(x, _folded_disc_weight(state, f"{prefix}.convs.{i}"), state[f"{prefix}.convs.{i}.bias"], stride=stride,
                     padding=pad, groups=groups)
        x = F.leaky_relu(x, 0.1)
        fmap.append(x)
    x = F.conv1d(x, _folded_disc_weight(state, f"{prefix}.conv_post"), state[f"{prefix}.conv_post.bias"], padding=1)
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def disc_conv(x, w, bias, stride, pad, groups, slope):
    """One discriminator layer on [B, C, L, W] (what nvse_disc_conv_forward_f32 computes): conv along L + leaky_relu."""
    y = F.conv2d(x, w.unsqueeze(-1), bias, stride=(stride, 1), padding=(pad, 0), groups=groups)
    return F.leaky_relu(y, slope) if slope != 1.0 else y
