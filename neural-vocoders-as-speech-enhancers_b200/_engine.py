"""Host side of the generator: owns the ``nvse_generator`` handle of one module,
folds weight-norm and uploads weights when they change, and runs the forward."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

_PRECISIONS = {"fp32": _lib.PRECISION_F32, "f32": _lib.PRECISION_F32, "float32": _lib.PRECISION_F32,
               "bf16": _lib.PRECISION_BF16, "bfloat16": _lib.PRECISION_BF16}
DEFAULT_PRECISION = "bf16"


def resolve_precision(name=None):
    name = name or os.environ.get("NVSE_B200_PRECISION") or DEFAULT_PRECISION
    try:
        return _PRECISIONS[str(name).lower()]
    except KeyError:
        raise ValueError(f"unknown precision {name!r}; use 'fp32' or 'bf16'") from None


def make_config(h, kind):
    """cfgs/*.json hyper-parameters (as read at hifigan.py:87-102, istftnet.py:275-297)
    -> struct nvse_generator_config."""
    cfg = _lib.GeneratorConfig()
    cfg.kind = kind
    cfg.in_channels = 80  # hard-coded in the reference (hifigan.py:89)
    cfg.initial_channel = int(h.upsample_initial_channel)
    rates, ksizes = list(h.upsample_rates), list(h.upsample_kernel_sizes)
    if len(rates) != len(ksizes) or not 1 <= len(rates) <= _lib.MAX_UPS:
        raise ValueError("upsample_rates / upsample_kernel_sizes mismatch")
    cfg.num_upsamples = len(rates)
    for i, (u, k) in enumerate(zip(rates, ksizes)):
        cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
    cfg.resblock_type = 1 if str(h.resblock) == "1" else 2
    rk, rd = list(h.resblock_kernel_sizes), list(h.resblock_dilation_sizes)
    if len(rk) != len(rd) or not 1 <= len(rk) <= _lib.MAX_KERNELS:
        raise ValueError("resblock_kernel_sizes / resblock_dilation_sizes mismatch")
    cfg.num_kernels = len(rk)
    for j, (k, ds) in enumerate(zip(rk, rd)):
        # ResBlock2 only ever builds two convolutions (hifigan.py:60: dilation=(1, 3) unpacked by index)
        ds = list(ds) if cfg.resblock_type == 1 else list(ds)[:2]
        if cfg.resblock_type == 1 and len(ds) != 3:
            raise ValueError("ResBlock1 uses exactly three dilations (hifigan.py:20-29)")
        cfg.resblock_kernel_sizes[j] = int(k)
        cfg.num_dilations[j] = len(ds)
        for m, d in enumerate(ds):
            cfg.resblock_dilations[j][m] = int(d)
    if kind == _lib.GEN_ISTFTNET:
        cfg.istft_n_fft, cfg.istft_hop = int(h.gen_istft_n_fft), int(h.gen_istft_hop_size)
    return cfg


class GeneratorEngine:
    """One per nn.Module instance; created lazily on the first forward."""

    def __init__(self, module, kind):
        self.kind = kind
        self.cfg = make_config(module.h, kind)
        self.handle = None
        self.device = None
        self.weights_key = None
        self.workspace = None

    def close(self):
        if self.handle is not None:
            try:
                _lib.load().nvse_generator_destroy(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()

    # ---- weights ---------------------------------------------------------------------
    @staticmethod
    def _conv_modules(module):
        for name, m in module.named_modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
                yield name, m

    def _key(self, module, dev):
        """Identity of the weights the handle holds: (data_ptr, version) of every parameter.  Walking
        ``named_parameters()`` costs ~0.5 ms per call (more than a batch-1 forward on the GPU), so the
        list of Parameter objects is cached; it is rebuilt when the module's structure changed
        (``remove_weight_norm`` / re-parametrisation swap Parameter objects: detected through the total
        parameter count of the conv modules) and, as a backstop, every 256 calls."""
        self._calls = getattr(self, "_calls", 0) + 1
        plist = getattr(self, "_plist", None)
        convs = getattr(self, "_convs", None)
        if convs is not None and (self._calls & 255) != 0:
            n = 0
            for m in convs:
                n += len(m._parameters)
            if n != self._nparam:
                plist = None
        else:
            plist = None
        if plist is None:
            self._convs = [m for _, m in self._conv_modules(module)]
            self._nparam = sum(len(m._parameters) for m in self._convs)
            plist = self._plist = list(module.parameters())
        return (dev.index,) + tuple([(p.data_ptr(), p._version) for p in plist])

    def _ensure(self, module, dev):
        lib = _lib.load()
        if self.handle is not None and self.device != dev:
            self.close()
        if self.handle is None:
            h = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(lib.nvse_generator_create(C.byref(self.cfg), C.byref(h)))
            self.handle, self.device, self.weights_key = h, dev, None
        key = self._key(module, dev)
        if key == self.weights_key:
            return
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        keep = []  # keep staging tensors alive until the copies are enqueued (same stream -> safe after)
        with torch.cuda.device(dev), torch.no_grad():
            for name, m in self._conv_modules(module):
                if hasattr(m, "weight_g") and hasattr(m, "weight_v"):  # still weight-normed (dim=0)
                    v = m.weight_v.detach().to(dev, torch.float32).contiguous()
                    g = m.weight_g.detach().to(dev, torch.float32).contiguous()
                    w = torch.empty_like(v)
                    _lib.check(lib.nvse_weight_norm_fold_f32(_lib.ptr(v), _lib.ptr(g), _lib.ptr(w), v.shape[0],
                                                             v[0].numel(), stream))
                    keep += [v, g]
                else:
                    w = m.weight.detach().to(dev, torch.float32).contiguous()
                b = m.bias.detach().to(dev, torch.float32).contiguous()
                for leaf, t in (("weight", w), ("bias", b)):
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.nvse_generator_set_weight(self.handle, f"{name}.{leaf}".encode(), _lib.ptr(t),
                                                             shape, t.dim(), stream))
                keep += [w, b]
            _lib.check(lib.nvse_generator_finalize(self.handle, stream))
            torch.cuda.current_stream(dev).synchronize()  # staging tensors may now be freed
        del keep
        self.weights_key = key

    # ---- forward ---------------------------------------------------------------------
    def forward(self, module, x, precision=None):
        if x.dim() != 3 or x.shape[1] != self.cfg.in_channels:
            raise RuntimeError(f"expected mel of shape [B, {self.cfg.in_channels}, frames], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and (x.requires_grad or (module.training and any(p.requires_grad for p in module.parameters()))):
            raise NotImplementedError(
                "the B200 generator implements inference only in this build; call it under torch.no_grad() "
                "(backward kernels are scheduled next, SURVEY.md §8f)")
        if x.is_cuda:
            dev = x.device
        else:
            if not torch.cuda.is_available():
                raise _lib.NvseError("the B200 generator needs a CUDA device: there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self._ensure(module, dev)
        prec = resolve_precision(precision or getattr(module, "precision", None))
        xd = x.detach().to(dev, torch.float32).contiguous()
        batch, _, frames = xd.shape
        n_out = lib.nvse_generator_out_samples(self.handle, frames)
        out = torch.empty((batch, n_out), dtype=torch.float32, device=dev)
        if batch == 0 or frames == 0:
            return out.to(x.device)
        need = lib.nvse_generator_workspace_bytes(self.handle, batch, frames, prec)
        if self.workspace is None or self.workspace.numel() < need or self.workspace.device != dev:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.nvse_generator_forward(self.handle, _lib.ptr(xd), batch, frames, _lib.ptr(out),
                                                  _lib.ptr(self.workspace), self.workspace.numel(), prec, stream))
        return out if x.is_cuda else out.to(x.device)
