"""Drop-ins for the reference's waveform discriminators (Models/models.py:15-113, 187-246) and the losses the
time-domain trainer takes from the same module (train_time_wi_inv.py:18-26, Models/models.py:604-669).

``MultiPeriodDiscriminator`` / ``MultiScaleDiscriminator`` keep the reference's constructor signatures, sub-module
tree and therefore its checkpoint keys (``discriminators.N.convs.M.weight_g`` / ``weight_v`` / ``bias``;
``weight_orig`` / ``weight_u`` / ``weight_v`` under spectral_norm) and random initialisation -- the layers ARE
``torch.nn.Conv1d`` / ``Conv2d`` sub-classes wrapped by torch's own ``weight_norm`` / ``spectral_norm``, only their
arithmetic is replaced: every convolution (with the following leaky_relu fused) runs in the sm_100a kernels of
csrc/disc.cu through ``nvse_disc_conv_forward_f32`` / ``nvse_disc_conv_backward_f32``, forward and backward, in the
reference's own channels-first layouts, so the feature maps come out exactly as the reference returns them.  The
re-parametrisations (a few element-wise ops per layer and spectral_norm's power iteration) stay in PyTorch.
There is no CPU path: inputs must live on a CUDA device."""
from __future__ import annotations

import ctypes as C
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    """utils.py:46-47."""
    return int((kernel_size * dilation - dilation) / 2)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _DiscConvFn(torch.autograd.Function):
    """y = leaky_relu(conv(x, w, stride, pad, groups) + bias, slope) over x [B, Cin, L, W] (convolution along L)."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, pad, groups, slope):
        if not x.is_cuda:
            raise _lib.NvseError("the B200 discriminators need CUDA tensors: there is no CPU fallback")
        lib = _lib.load()
        x = x.contiguous().float()
        w = weight.detach().contiguous().float()
        b = None if bias is None else bias.detach().contiguous().float()
        B, Cin, L, W = x.shape
        Cout = w.shape[0]
        Lo = int(lib.nvse_disc_conv_out_len(L, k, stride, pad))
        if Lo < 1:
            raise _lib.NvseError(f"discriminator conv: input of {L} rows is shorter than the kernel ({k}, padding {pad})")
        y = torch.empty((B, Cout, Lo, W), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.nvse_disc_conv_forward_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), B, Cin, Cout, L, W, k,
                                                      stride, pad, groups, float(slope), _stream()))
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (k, stride, pad, groups, float(slope), bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        k, stride, pad, groups, slope, has_bias = ctx.cfg
        lib = _lib.load()
        B, Cin, L, W = x.shape
        Cout = w.shape[0]
        dy = dy.contiguous().float()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty((Cout,), device=x.device, dtype=torch.float32) if need_b else None
        nbytes = int(lib.nvse_disc_conv_backward_scratch_bytes(B, Cin, Cout, L, W, k, stride, pad, groups))
        scratch = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device):
            _lib.check(lib.nvse_disc_conv_backward_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(dw),
                                                       _lib.ptr(db), B, Cin, Cout, L, W, k, stride, pad, groups, slope,
                                                       _lib.ptr(scratch), nbytes, _stream()))
        return dx, dw, db, None, None, None, None, None


class _AvgPoolFn(torch.autograd.Function):
    """AvgPool1d(k, stride, padding) over the last axis (MultiScaleDiscriminator.meanpools, Models/models.py:225-228)."""

    @staticmethod
    def forward(ctx, x, k, stride, pad):
        if not x.is_cuda:
            raise _lib.NvseError("the B200 discriminators need CUDA tensors: there is no CPU fallback")
        lib = _lib.load()
        x = x.contiguous().float()
        T = x.shape[-1]
        rows = x.numel() // T
        To = (T + 2 * pad - k) // stride + 1
        y = torch.empty(x.shape[:-1] + (To,), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.nvse_avgpool1d_f32(_lib.ptr(x), _lib.ptr(y), rows, T, k, stride, pad, _stream()))
        ctx.cfg = (tuple(x.shape), k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, k, stride, pad = ctx.cfg
        lib = _lib.load()
        dy = dy.contiguous().float()
        dx = torch.empty(shape, device=dy.device, dtype=torch.float32)
        T = shape[-1]
        with torch.cuda.device(dy.device):
            _lib.check(lib.nvse_avgpool1d_backward_f32(_lib.ptr(dy), _lib.ptr(dx), dx.numel() // T, T, k, stride, pad, _stream()))
        return dx, None, None, None


class _DiscConv1d(nn.Conv1d):
    """``nn.Conv1d`` whose arithmetic is the B200 kernel; ``slope`` fuses the leaky_relu that follows it."""

    def forward(self, x, slope=1.0):
        y = _DiscConvFn.apply(x.unsqueeze(-1), self.weight, self.bias, self.kernel_size[0], self.stride[0], self.padding[0],
                              self.groups, slope)
        return y.squeeze(-1)


class _DiscConv2d(nn.Conv2d):
    """``nn.Conv2d`` with a (k, 1) kernel, (stride, 1) stride and (pad, 0) padding -- the only kind DiscriminatorP uses."""

    def forward(self, x, slope=1.0):
        if self.kernel_size[1] != 1 or self.stride[1] != 1 or self.padding[1] != 0 or self.groups != 1:
            raise _lib.NvseError("the B200 period discriminator supports (k, 1) kernels with (s, 1) strides only")
        return _DiscConvFn.apply(x, self.weight.squeeze(-1), self.bias, self.kernel_size[0], self.stride[0], self.padding[0], 1, slope)


def _weight_norm(m):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        return nn.utils.weight_norm(m)  # old-style keys (weight_g / weight_v), like Models/models.py:5


class DiscriminatorP(nn.Module):
    """Models/models.py:15-87."""

    def __init__(self, period, kernel_size=5, stride=3, use_spectral_norm=False):
        super().__init__()
        self.period = period
        self.use_spectral_norm = use_spectral_norm
        norm_f = nn.utils.spectral_norm if use_spectral_norm else _weight_norm
        pad = (get_padding(5, 1), 0)
        self.convs = nn.ModuleList([
            norm_f(_DiscConv2d(1, 32, (kernel_size, 1), (stride, 1), padding=pad)),
            norm_f(_DiscConv2d(32, 128, (kernel_size, 1), (stride, 1), padding=pad)),
            norm_f(_DiscConv2d(128, 512, (kernel_size, 1), (stride, 1), padding=pad)),
            norm_f(_DiscConv2d(512, 1024, (kernel_size, 1), (stride, 1), padding=pad)),
            norm_f(_DiscConv2d(1024, 1024, (kernel_size, 1), 1, padding=(2, 0))),
        ])
        self.conv_post = norm_f(_DiscConv2d(1024, 1, (3, 1), 1, padding=(1, 0)))

    def forward(self, x):
        fmap = []
        if x.ndim == 2:
            x = x.unsqueeze(1)
        b, c, t = x.shape
        if t % self.period != 0:
            n_pad = self.period - (t % self.period)
            x = F.pad(x, (0, n_pad), "reflect")
            t = t + n_pad
        x = x.reshape(b, c, t // self.period, self.period)
        for conv in self.convs:
            x = conv(x, LRELU_SLOPE)
            fmap.append(x)
        x = self.conv_post(x)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class MultiPeriodDiscriminator(nn.Module):
    """Models/models.py:90-113 (the fifth discriminator takes ``mpd_reshapes[-1]``, as there)."""

    def __init__(self, mpd_reshapes):
        super().__init__()
        self.discriminators = nn.ModuleList([
            DiscriminatorP(mpd_reshapes[0]), DiscriminatorP(mpd_reshapes[1]), DiscriminatorP(mpd_reshapes[2]),
            DiscriminatorP(mpd_reshapes[3]), DiscriminatorP(mpd_reshapes[-1]),
        ])

    def forward(self, y, y_hat):
        return _run_pairs(self.discriminators, lambda i, v: v, y, y_hat)


class DiscriminatorS(nn.Module):
    """Models/models.py:187-214."""

    def __init__(self, use_spectral_norm=False):
        super().__init__()
        self.use_spectral_norm = use_spectral_norm
        norm_f = nn.utils.spectral_norm if use_spectral_norm else _weight_norm
        self.convs = nn.ModuleList([
            norm_f(_DiscConv1d(1, 128, 15, 1, padding=7)),
            norm_f(_DiscConv1d(128, 128, 41, 2, groups=4, padding=20)),
            norm_f(_DiscConv1d(128, 256, 41, 2, groups=16, padding=20)),
            norm_f(_DiscConv1d(256, 512, 41, 4, groups=16, padding=20)),
            norm_f(_DiscConv1d(512, 1024, 41, 4, groups=16, padding=20)),
            norm_f(_DiscConv1d(1024, 1024, 41, 1, groups=16, padding=20)),
            norm_f(_DiscConv1d(1024, 1024, 5, 1, padding=2)),
        ])
        self.conv_post = norm_f(_DiscConv1d(1024, 1, 3, 1, padding=1))

    def forward(self, x):
        fmap = []
        if x.ndim == 2:
            x = x.unsqueeze(1)
        for conv in self.convs:
            x = conv(x, LRELU_SLOPE)
            fmap.append(x)
        x = self.conv_post(x)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class _MeanPool(nn.Module):
    """``AvgPool1d(4, 2, padding=2)`` (parameter-free, so the state dict is unchanged)."""

    def __init__(self, kernel_size, stride, padding):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return _AvgPoolFn.apply(x, self.kernel_size, self.stride, self.padding)

    def extra_repr(self):
        return f"kernel_size=({self.kernel_size},), stride=({self.stride},), padding=({self.padding},)"


class MultiScaleDiscriminator(nn.Module):
    """Models/models.py:217-246: the first scale under spectral_norm, each further scale on the mean-pooled signals."""

    def __init__(self):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorS(use_spectral_norm=True), DiscriminatorS(), DiscriminatorS()])
        self.meanpools = nn.ModuleList([_MeanPool(4, 2, 2), _MeanPool(4, 2, 2)])

    def forward(self, y, y_hat):
        if y.ndim == 2:  # the reference's AvgPool1d takes [C, T] as an unbatched signal; make the batch axis explicit
            y, y_hat = y.unsqueeze(1), y_hat.unsqueeze(1)
        return _run_pairs(self.discriminators, lambda i, v: v if i == 0 else self.meanpools[i - 1](v), y, y_hat, chain=True)


def _run_pairs(discriminators, prepare, y, y_hat, chain=False):
    """The common loop of both ``forward``s (Models/models.py:103-113, 231-246): every sub-discriminator sees the real and
    the generated batch -- in that order, because spectral_norm advances its power iteration on every call."""
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    for i, d in enumerate(discriminators):
        yi, yhi = prepare(i, y), prepare(i, y_hat)
        if chain:
            y, y_hat = yi, yhi
        if yi.shape == yhi.shape and not (d.training and getattr(d, "use_spectral_norm", False)):
            # one pass over the concatenated batch: every sample is independent and both calls would use the same weights
            # (a training-mode spectral_norm layer re-estimates its weight on every call, so that case keeps the two calls)
            n = yi.shape[0]
            out, fmap = d(torch.cat([yi, yhi], dim=0))
            y_d_r, y_d_g = out[:n], out[n:]
            fmap_r, fmap_g = [f[:n] for f in fmap], [f[n:] for f in fmap]
        else:
            y_d_r, fmap_r = d(yi)
            y_d_g, fmap_g = d(yhi)
        y_d_rs.append(y_d_r)
        fmap_rs.append(fmap_r)
        y_d_gs.append(y_d_g)
        fmap_gs.append(fmap_g)
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


# ---- losses of the time-domain trainer (plain tensor arithmetic on the discriminator outputs) -------------------------
def feature_loss(fmap_r, fmap_g):
    """Models/models.py:604-610."""
    loss = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss = loss + torch.mean(torch.abs(rl - gl))
    return loss


def _disc_loss(real_outputs, generated_outputs, real_term, generated_term):
    loss, r_losses, g_losses = 0, [], []
    for dr, dg in zip(real_outputs, generated_outputs):
        r_loss, g_loss = real_term(dr), generated_term(dg)
        loss = loss + (r_loss + g_loss)
        r_losses.append(r_loss.item())
        g_losses.append(g_loss.item())
    return loss, r_losses, g_losses


def hinge_discriminator_loss(disc_real_outputs, disc_generated_outputs):
    """Models/models.py:613-625."""
    return _disc_loss(disc_real_outputs, disc_generated_outputs, lambda dr: torch.mean(torch.clamp(1 - dr, min=0)),
                      lambda dg: torch.mean(torch.clamp(1 + dg, min=0)))


def ls_discriminator_loss(disc_real_outputs, disc_generated_outputs):
    """Models/models.py:627-639."""
    return _disc_loss(disc_real_outputs, disc_generated_outputs, lambda dr: torch.mean((1 - dr) ** 2), lambda dg: torch.mean(dg ** 2))


def _gen_loss(disc_outputs, term):
    loss, gen_losses = 0, []
    for dg in disc_outputs:
        l = term(dg)
        gen_losses.append(l)
        loss = loss + l
    return loss, gen_losses


def hinge_generator_loss(disc_outputs):
    """Models/models.py:642-650."""
    return _gen_loss(disc_outputs, lambda dg: torch.mean(torch.clamp(1 - dg, min=0)))


def ls_generator_loss(disc_outputs):
    """Models/models.py:652-660."""
    return _gen_loss(disc_outputs, lambda dg: torch.mean((1 - dg) ** 2))
