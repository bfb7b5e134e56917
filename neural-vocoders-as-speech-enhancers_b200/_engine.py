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


def resolve_train_precision(module):
    """Precision of the TRAINING path (forward with a tape + CUDA backward).  The reference trains in fp32
    (train_time_wi_inv.py has no autocast), so an unmodified training script gets the fp32 kernels: the tensor-core
    backward (bf16 operands; per-tensor gradients 3-7.5 % relative L2 off fp32) is an explicit opt-in --
    ``module.train_precision = "bf16"``, ``NVSE_B200_TRAIN_PRECISION=bf16``, or a ``module.precision`` the caller set
    by hand.  The inference switch ``NVSE_B200_PRECISION`` does not touch training."""
    name = (getattr(module, "train_precision", None) or os.environ.get("NVSE_B200_TRAIN_PRECISION")
            or getattr(module, "precision", None) or "fp32")
    try:
        return _PRECISIONS[str(name).lower()]
    except KeyError:
        raise ValueError(f"unknown training precision {name!r}; use 'fp32' or 'bf16'") from None


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


def _with_index(dev):
    """``torch.device("cuda")`` and ``torch.device("cuda", 0)`` compare unequal: every device the engine keys its native handle
    on carries an explicit index (otherwise a caller mixing the two forms would rebuild the handle on every call)."""
    dev = torch.device(dev)
    if dev.type == "cuda" and dev.index is None:
        return torch.device("cuda", torch.cuda.current_device())
    return dev


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

    def invalidate(self):
        """Forget which weights the handle holds (and the cached list of conv modules): the next forward re-reads every
        parameter.  Called by the module after ``load_state_dict``, ``remove_weight_norm`` and ``.to()``; call it by hand
        after writing through ``p.data`` or swapping a sub-module, which neither autograd's version counters nor the
        hooks can see."""
        self.weights_key = None
        self._convs = None
        self._batched = None

    # ---- weights ---------------------------------------------------------------------
    @staticmethod
    def _conv_modules(module):
        for name, m in module.named_modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
                yield name, m

    def _key(self, module, dev):
        """Identity of the weights the handle holds: (data_ptr, version) of every parameter of every conv module, read
        from the modules' CURRENT ``_parameters`` (so in-place updates, ``remove_weight_norm`` and re-assigned Parameter
        objects are all seen).  ``named_parameters()`` over the module tree costs ~0.5 ms per call (more than a batch-1
        forward on the GPU), so the list of conv modules is cached and only re-collected every 64 calls (a sub-module
        swapped for a new one is the one change this can miss, for at most 64 calls; writes through ``p.data`` bypass
        autograd's version counter by design and are not seen either -- call ``invalidate()`` after either)."""
        self._calls = getattr(self, "_calls", 0) + 1
        convs = getattr(self, "_convs", None)
        if convs is None or (self._calls & 63) == 0:
            convs = self._convs = [m for _, m in self._conv_modules(module)]
        key = [dev.index]
        for m in convs:
            for p in m._parameters.values():
                if p is not None:
                    key.append(p.data_ptr())
                    key.append(p._version)
        return tuple(key)

    def _layer_names(self, lib):
        names = getattr(self, "_names", None)
        if names is None:
            buf = C.create_string_buffer(256)
            names = []
            for i in range(lib.nvse_generator_num_layers(self.handle)):
                _lib.check(lib.nvse_generator_layer_name(self.handle, i, buf, len(buf)))
                names.append(buf.value.decode())
            self._names = names
        return names

    def _load_batched(self, module, dev, train, lib, stream):
        """All layers in two launches (nvse_generator_load_weights) when every parameter already is a contiguous fp32
        CUDA tensor on ``dev`` -- the training case, where the weights change every step.  Returns False if not eligible."""
        mods = dict(self._conv_modules(module))
        names = self._layer_names(lib)
        n = len(names)
        w_arr, g_arr, b_arr = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
        for i, name in enumerate(names):
            m = mods.get(name)
            if m is None:
                return False
            wn = hasattr(m, "weight_g") and hasattr(m, "weight_v")
            tensors = (m.weight_v, m.weight_g, m.bias) if wn else (m.weight, m.bias)
            for t in tensors:
                if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                    return False
            w_arr[i] = (m.weight_v if wn else m.weight).data_ptr()
            g_arr[i] = m.weight_g.data_ptr() if wn else None
            b_arr[i] = m.bias.data_ptr()
        _lib.check(lib.nvse_generator_load_weights(self.handle, w_arr, g_arr, b_arr, n, 1 if train else 0, stream))
        self._batched = (w_arr, g_arr, [mods[name] for name in names])
        return True

    def _ensure(self, module, dev, train=False):
        lib = _lib.load()
        dev = _with_index(dev)
        if self.handle is not None and self.device != dev:
            self.close()
        if self.handle is None:
            h = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(lib.nvse_generator_create(C.byref(self.cfg), C.byref(h)))
            self.handle, self.device, self.weights_key = h, dev, None
            self._names = None
        key = self._key(module, dev)
        if key == self.weights_key and (not train or getattr(self, "_train_loaded", False)):
            return
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev), torch.no_grad():
            if self._load_batched(module, dev, train, lib, stream):
                self.weights_key, self._train_loaded = key, train
                return
        self._batched = None
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
        self.weights_key, self._train_loaded = key, True  # the per-layer path builds the backward's copies lazily

    # ---- forward ---------------------------------------------------------------------
    def forward(self, module, x, precision=None, pcm16=False, out=None, frames=None):
        """``pcm16=True`` (inference only): the waveform as int16 PCM, quantised like ``sf.write(..., 'PCM_16')``
        (infers/inference_hifigan.py:93) inside the last kernel instead of float32.  ``out`` (inference only): a
        contiguous device tensor ``[B, samples]`` of the result's dtype to write into instead of allocating one.
        ``frames`` (inference only, the 16-bit tensor-core path of either generator): int tensor ``[B]``, the mel frames of each utterance of a
        batch padded to the longest one -- utterance ``b`` comes out bit-identical to passing ``x[b:b+1, :, :frames[b]]``
        alone (``nvse_generator_forward_ragged``); its samples beyond its own length are undefined."""
        if x.dim() != 3 or x.shape[1] != self.cfg.in_channels:
            raise RuntimeError(f"expected mel of shape [B, {self.cfg.in_channels}, frames], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and (x.requires_grad or (module.training and any(p.requires_grad for p in module.parameters()))):
            # training step (train_time_wi_inv.py:173-236): forward with a tape + the CUDA backward, fp32 unless the
            # caller opted in to the tensor-core training path (resolve_train_precision)
            if pcm16 or frames is not None:
                raise RuntimeError("PCM_16 output / ragged batches are inference features: call them under torch.no_grad() / module.eval()")
            return _GeneratorTrainFn.apply(self, module, x, *[p for _, p in module.named_parameters()])
        if x.is_cuda:
            dev = x.device
        else:
            if not torch.cuda.is_available():
                raise _lib.NvseError("the B200 generator needs a CUDA device: there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self._ensure(module, dev)
        prec = resolve_precision(precision or getattr(module, "precision", None))
        per_item = frames
        xd = x.detach().to(dev, torch.float32).contiguous()
        batch, _, frames = xd.shape
        n_out = lib.nvse_generator_out_samples(self.handle, frames)
        odt = torch.int16 if pcm16 else torch.float32
        if out is not None:
            if not (out.is_cuda and out.device == dev and out.dtype == odt and tuple(out.shape) == (batch, n_out) and out.is_contiguous()):
                raise RuntimeError(f"out must be a contiguous {odt} tensor of shape {(batch, n_out)} on {dev}")
        else:
            out = torch.empty((batch, n_out), dtype=odt, device=dev)
        if batch == 0 or frames == 0:
            return out.to(x.device)
        need = lib.nvse_generator_workspace_bytes(self.handle, batch, frames, prec)
        if self.workspace is None or self.workspace.numel() < need or self.workspace.device != dev:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if per_item is not None:
                if pcm16:
                    raise RuntimeError("ragged batches deliver float output (quantise with Vocoder.pcm16)")
                fd = torch.as_tensor(per_item).to(dev, torch.int32).contiguous()
                if fd.shape != (batch,) or (batch and (int(fd.min()) < 1 or int(fd.max()) > frames)):
                    raise RuntimeError(f"frames must be an int tensor [{batch}] with values in [1, {frames}]")
                _lib.check(lib.nvse_generator_forward_ragged(self.handle, _lib.ptr(xd), batch, frames, _lib.ptr(fd), _lib.ptr(out),
                                                             _lib.ptr(self.workspace), self.workspace.numel(), prec, stream))
            else:
                fwd = lib.nvse_generator_forward_pcm16 if pcm16 else lib.nvse_generator_forward
                _lib.check(fwd(self.handle, _lib.ptr(xd), batch, frames, _lib.ptr(out),
                               _lib.ptr(self.workspace), self.workspace.numel(), prec, stream))
        return out if x.is_cuda else out.to(x.device)

    # ---- fused wav -> wav (inference, 16-bit path) -------------------------------------------
    def can_vocode(self, module, dev):
        """True when ``vocode`` applies: 16-bit tensor-core precision and a conv_pre the tensor-core launch takes."""
        if resolve_precision(getattr(module, "precision", None)) != _lib.PRECISION_BF16:
            return False
        if self.handle is None or self.device != _with_index(dev) or self.weights_key is None:
            self._ensure(module, dev)  # the pitch is a property of the configuration: a loaded handle answers without a re-check
        return int(_lib.load().nvse_vocoder_mel_pitch(self.handle)) > 0

    def vocode(self, module, frontend, wav, lengths=None, pcm16=False, out=None):
        """``nvse_vocoder_forward``: wav ``[B, T]`` on the device -> waveform ``[B, samples]`` (int16 PCM with ``pcm16``) in one
        library call -- the front-end writes the log-mel straight into conv_pre's staging layout (no ``[B, 80, F]`` tensor, no
        transpose pass); bit-identical to ``generator(mel_spectrogram(wav))``.  ``lengths``: int tensor ``[B]`` of samples per
        utterance of a padded batch (float output only)."""
        lib = _lib.load()
        dev = wav.device
        self._ensure(module, dev)
        wd = wav.detach().to(torch.float32)
        if wd.dim() != 2 or wd.stride(-1) != 1:
            wd = wd.reshape(wd.shape[0], -1).contiguous()
        batch, samples = wd.shape
        frames = int(lib.nvse_frontend_num_frames(frontend, samples))
        n_out = int(lib.nvse_generator_out_samples(self.handle, frames))
        odt = torch.int16 if pcm16 else torch.float32
        if out is not None:
            if not (out.is_cuda and out.device == dev and out.dtype == odt and tuple(out.shape) == (batch, n_out) and out.is_contiguous()):
                raise RuntimeError(f"out must be a contiguous {odt} tensor of shape {(batch, n_out)} on {dev}")
        else:
            out = torch.empty((batch, n_out), dtype=odt, device=dev)
        if batch == 0:
            return out
        need = int(lib.nvse_vocoder_workspace_bytes(frontend, self.handle, batch, samples))
        if need == 0:
            raise _lib.NvseError("the fused wav -> wav call does not apply to this generator (conv_pre is not on the tensor cores)")
        if self.workspace is None or self.workspace.numel() < need or self.workspace.device != dev:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        ld = None
        if lengths is not None:
            if pcm16:
                raise RuntimeError("ragged batches deliver float output (quantise with Vocoder.pcm16)")
            ld = torch.as_tensor(lengths).to(dev, torch.int32).contiguous()
            if ld.shape != (batch,):
                raise RuntimeError(f"lengths must have shape [{batch}], got {tuple(ld.shape)}")
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.nvse_vocoder_forward(frontend, self.handle, _lib.ptr(wd), batch, samples, wd.stride(0) if batch > 1 else samples,
                                                _lib.ptr(ld), None if pcm16 else _lib.ptr(out), _lib.ptr(out) if pcm16 else None,
                                                _lib.ptr(self.workspace), self.workspace.numel(), stream))
        return out

    # ---- training ---------------------------------------------------------------------
    def forward_train(self, module, x):
        """Forward that keeps every convolution input on an fp32 tape; returns (out, tape, precision).  The arithmetic is
        fp32 CUDA cores by default, tensor cores (bf16 operands, fp32 accumulate / tape / master weights) when opted in
        (resolve_train_precision)."""
        if x.dim() != 3 or x.shape[1] != self.cfg.in_channels:
            raise RuntimeError(f"expected mel of shape [B, {self.cfg.in_channels}, frames], got {tuple(x.shape)}")
        if not x.is_cuda and not torch.cuda.is_available():
            raise _lib.NvseError("the B200 generator needs a CUDA device: there is no CPU fallback")
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self._ensure(module, dev, train=True)
        xd = x.detach().to(dev, torch.float32).contiguous()
        batch, _, frames = xd.shape
        if batch == 0 or frames == 0:
            raise RuntimeError("empty batch on the training path")
        out = torch.empty((batch, lib.nvse_generator_out_samples(self.handle, frames)), dtype=torch.float32, device=dev)
        tape = torch.empty(lib.nvse_generator_tape_bytes(self.handle, batch, frames), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            prec = resolve_train_precision(module)
            _lib.check(lib.nvse_generator_forward_train(self.handle, _lib.ptr(xd), batch, frames, _lib.ptr(out),
                                                        _lib.ptr(tape), tape.numel(), prec, stream))
        return out, tape, prec

    def backward(self, module, out, dout, tape, frames, want_dmel, precision):
        """-> (dict: folded tensor name -> gradient view, dmel or None)"""
        lib = _lib.load()
        dev = out.device
        batch = out.shape[0]
        grads = torch.empty(lib.nvse_generator_grad_elems(self.handle), dtype=torch.float32, device=dev)
        dmel = torch.empty((batch, self.cfg.in_channels, frames), dtype=torch.float32, device=dev) if want_dmel else None
        need = lib.nvse_generator_backward_workspace_bytes(self.handle, batch, frames)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        dout = dout.to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.nvse_generator_backward(self.handle, batch, frames, _lib.ptr(out), _lib.ptr(dout), _lib.ptr(tape),
                                                   tape.numel(), _lib.ptr(grads), _lib.ptr(dmel), _lib.ptr(ws), ws.numel(),
                                                   precision, stream))
        self._grads_flat = grads
        views = {}
        off, n = C.c_int64(), C.c_int64()
        for name, m in self._conv_modules(module):
            for leaf, shape in (("weight", tuple(m.weight.shape)), ("bias", tuple(m.bias.shape))):
                _lib.check(lib.nvse_generator_grad_offset(self.handle, f"{name}.{leaf}".encode(), C.byref(off), C.byref(n)))
                views[f"{name}.{leaf}"] = grads[off.value:off.value + n.value].view(shape)
        return views, dmel


class _GeneratorTrainFn(torch.autograd.Function):
    """HiFiGAN.forward / iSTFTNet.forward as one autograd node: ``apply(engine, module, mel, *module.named_parameters())``."""

    @staticmethod
    def forward(ctx, engine, module, x, *params):
        out, tape, prec = engine.forward_train(module, x)
        ctx.engine, ctx.module, ctx.tape, ctx.prec = engine, module, tape, prec
        ctx.frames, ctx.x_device, ctx.x_dtype = x.shape[2], x.device, x.dtype
        ctx.key = engine.weights_key
        ctx.save_for_backward(out)
        return out if x.is_cuda else out.to(x.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        engine, module = ctx.engine, ctx.module
        (out,) = ctx.saved_tensors
        if engine.weights_key != ctx.key or engine._key(module, out.device) != ctx.key:
            raise RuntimeError("generator parameters changed between forward and backward")
        lib = _lib.load()
        want_dmel = ctx.needs_input_grad[2]
        folded, dmel = engine.backward(module, out, dout, ctx.tape, ctx.frames, want_dmel, ctx.prec)
        ctx.tape = None
        grads = {}
        dev = out.device
        batched = getattr(engine, "_batched", None)
        with torch.cuda.device(dev), torch.no_grad():
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if batched is not None:  # parameters live on the GPU: weight_norm backward of all layers in one launch
                w_arr, g_arr, mods = batched
                flat = engine._grads_flat
                dv_flat = torch.empty_like(flat)
                dg_flat = torch.empty(lib.nvse_generator_total_rows(engine.handle), dtype=torch.float32, device=dev)
                _lib.check(lib.nvse_generator_weight_norm_backward(engine.handle, w_arr, g_arr, _lib.ptr(flat), _lib.ptr(dv_flat),
                                                                   _lib.ptr(dg_flat), len(mods), stream))
                row = 0
                for name, m in zip(engine._layer_names(lib), mods):
                    grads[f"{name}.bias"] = folded[f"{name}.bias"]
                    dw = folded[f"{name}.weight"]
                    rows = m.weight_v.shape[0] if hasattr(m, "weight_v") else m.weight.shape[0]
                    if hasattr(m, "weight_g") and hasattr(m, "weight_v"):
                        off = dw.storage_offset()
                        grads[f"{name}.weight_v"] = dv_flat[off:off + dw.numel()].view(dw.shape)
                        grads[f"{name}.weight_g"] = dg_flat[row:row + rows].view(m.weight_g.shape)
                    else:
                        grads[f"{name}.weight"] = dw
                    row += rows
            for name, m in (engine._conv_modules(module) if batched is None else ()):
                grads[f"{name}.bias"] = folded[f"{name}.bias"]
                dw = folded[f"{name}.weight"]
                if hasattr(m, "weight_g") and hasattr(m, "weight_v"):  # weight_norm(dim=0): w = g * v / ||v||
                    v = m.weight_v.detach().to(dev, torch.float32).contiguous()
                    g = m.weight_g.detach().to(dev, torch.float32).contiguous()
                    dv, dg = torch.empty_like(v), torch.empty_like(g)
                    _lib.check(lib.nvse_weight_norm_backward_f32(_lib.ptr(v), _lib.ptr(g), _lib.ptr(dw), _lib.ptr(dv), _lib.ptr(dg),
                                                                 v.shape[0], v[0].numel(), stream))
                    grads[f"{name}.weight_v"], grads[f"{name}.weight_g"] = dv, dg
                else:
                    grads[f"{name}.weight"] = dw
        ordered = []
        for i, (name, p) in enumerate(module.named_parameters()):
            gr = grads.get(name) if ctx.needs_input_grad[3 + i] else None
            ordered.append(None if gr is None else gr.to(p.device, p.dtype))
        if dmel is not None:
            dmel = dmel.to(ctx.x_device, ctx.x_dtype)
        return (None, None, dmel, *ordered)
