"""Batched wav -> mel -> generator -> wav, the loop body of the reference's inference script
(infers/inference_hifigan.py:67-95: ``x = get_mel(wav); y = generator(x)``) for a whole list of
equal-length utterances, from HOST buffers to HOST buffers.

The reference loops one utterance at a time on one device; here a shard of utterances is pushed
through in micro-batches: pinned host -> device copy, fused front-end kernel, generator kernels,
device -> pinned host copy, all stream-ordered with no synchronisation inside the loop."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .dataset import mel_spectrogram, frontend_handle
from ._engine import GeneratorEngine
from .shard import bucket_by_length, bucket_padded


class Vocoder:
    def __init__(self, generator, h, micro_batch=32, device=None):
        self.generator = generator
        self.h = h
        self.micro_batch = int(micro_batch)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type == "cuda" and self.device.index is None:  # "cuda" != "cuda:0" for torch: always carry the index
            self.device = torch.device("cuda", torch.cuda.current_device())

    def mel(self, wav_dev):
        h = self.h
        return mel_spectrogram(wav_dev, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax)

    def _engine(self):
        gen = self.generator
        if gen._engine is None:
            object.__setattr__(gen, "_engine", GeneratorEngine(gen, gen._kind))
        return gen._engine

    def fused(self):
        """True when wav -> mel -> generator runs as ONE library call (``nvse_vocoder_forward``: the front-end writes the log-mel
        straight into conv_pre's tensor-core staging layout): the 16-bit path of a generator whose conv_pre the tensor-core
        launch takes.  Otherwise the two stages are called one after the other -- the results are bit-identical."""
        gen = self.generator
        if gen is None or not hasattr(gen, "_kind") or gen.training or getattr(self, "no_fuse", False):
            return False
        return self._engine().can_vocode(gen, self.device)

    def vocode(self, wav_dev, lengths=None, pcm16=False, out=None):
        """One micro-batch on the device: wav ``[B, T]`` -> waveform ``[B, samples]`` (fused call where it applies)."""
        h = self.h
        if self.fused() and wav_dev.shape[-1] > int(h.n_fft) // 2:
            fe = frontend_handle(h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax, self.device)
            return self._engine().vocode(self.generator, fe, wav_dev, lengths=lengths, pcm16=pcm16, out=out)
        if lengths is not None:
            n = torch.as_tensor(lengths).to(self.device, torch.int32)
            mel = mel_spectrogram(wav_dev, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax, lengths=n)
            frames = (1 + torch.div(n, int(h.hop_size), rounding_mode="floor")).to(torch.int32)
            return self.generator(mel, frames=frames)
        mel = self.mel(wav_dev)
        return self.generator.forward_pcm16(mel, out=out) if pcm16 else self.generator(mel, out=out)

    @torch.no_grad()
    def run_device(self, wav_dev, out_dev=None):
        """Device-resident variant: wav_dev [U, T] on the GPU -> [U, T_out] on the GPU."""
        outs = []
        for s in range(0, wav_dev.shape[0], self.micro_batch):
            y = self.vocode(wav_dev[s:s + self.micro_batch])
            if out_dev is not None:
                out_dev[s:s + y.shape[0]].copy_(y)
            else:
                outs.append(y)
        return out_dev if out_dev is not None else torch.cat(outs, 0)

    def pcm16(self, y):
        """Device float waveform -> int16 PCM as the reference's sf.write(..., 'PCM_16') stores it
        (infers/inference_hifigan.py:93): round(x * 32767), clipped."""
        y = y.contiguous()
        out = torch.empty(y.shape, dtype=torch.int16, device=y.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        _lib.check(_lib.load().nvse_pcm16_from_f32(_lib.ptr(y), _lib.ptr(out), y.numel(), st))
        return out

    @torch.no_grad()
    def run_host(self, wav_host, out_host=None, pcm16=False):
        """wav_host: [U, T] float32 CPU tensor (pinned for async copies).  Returns (and fills, if
        given) a CPU tensor [U, T_out] (int16 PCM with ``pcm16=True``).  Copies run on two side streams so that the host -> device copy
        of micro-batch i+1 and the device -> host copy of micro-batch i-1 overlap the kernels of
        micro-batch i; the caller synchronises the current stream (all side-stream work is joined to it)."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._stage = None
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        starts = list(range(0, wav_host.shape[0], self.micro_batch))
        # Two device input and two device output buffers, kept across calls and used alternately: no allocation (and no
        # allocator-driven cudaMalloc while blocks wait for their cross-stream use to end) on the steady-state path.
        # consumed[k]: the kernels have read input buffer k; drained[k]: output buffer k has reached the host.
        key = (int(wav_host.shape[-1]), bool(pcm16), min(self.micro_batch, max(1, wav_host.shape[0])))
        if starts and (self._stage is None or self._stage["key"] != key):
            self._stage = {"key": key, "in": [torch.empty((key[2], key[0]), dtype=torch.float32, device=dev) for _ in range(2)],
                           "out": [None, None], "consumed": [None, None], "drained": [None, None]}
        stg = self._stage

        def upload(i):
            s, k = starts[i], i & 1
            n = min(self.micro_batch, wav_host.shape[0] - s)
            with torch.cuda.stream(self._h2d):
                if stg["consumed"][k] is not None:
                    self._h2d.wait_event(stg["consumed"][k])
                chunk = stg["in"][k][:n]
                chunk.copy_(wav_host[s:s + n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._h2d)
            return chunk, ev

        nxt = upload(0) if starts else None
        for i, s in enumerate(starts):
            chunk, ev = nxt
            k = i & 1
            nxt = upload(i + 1) if i + 1 < len(starts) else None
            cur.wait_event(ev)
            if stg["drained"][k] is not None:
                cur.wait_event(stg["drained"][k])
            obuf = stg["out"][k]
            obuf = obuf[:chunk.shape[0]] if obuf is not None and obuf.shape[0] >= chunk.shape[0] else None
            # quantised inside the generator's last kernel with pcm16: half the bytes cross PCIe, no extra pass
            y = self.vocode(chunk, pcm16=pcm16, out=obuf)
            stg["consumed"][k] = torch.cuda.Event()
            stg["consumed"][k].record(cur)
            if stg["out"][k] is None or stg["out"][k].shape[0] < y.shape[0]:
                stg["out"][k] = y
            y = y.reshape(y.shape[0], -1)
            if out_host is None:
                out_host = torch.empty((wav_host.shape[0], y.shape[1]), dtype=y.dtype, pin_memory=True)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done)
                out_host[s:s + y.shape[0]].copy_(y, non_blocking=True)
                stg["drained"][k] = torch.cuda.Event()
                stg["drained"][k].record(self._d2h)
        cur.wait_stream(self._d2h)
        if out_host is None:  # no utterances: an empty result of the right width, like the device-resident path
            h = self.h
            n_out = 1
            for u in h.upsample_rates:
                n_out *= int(u)
            if hasattr(h, "gen_istft_hop_size"):
                n_out *= int(h.gen_istft_hop_size)
            frames = 1 + wav_host.shape[-1] // int(h.hop_size) if wav_host.dim() == 2 else 0
            out_host = torch.empty((0, frames * n_out), dtype=torch.int16 if pcm16 else torch.float32)
        return out_host

    @torch.no_grad()
    def run_list(self, wavs, max_pad=0.25):
        """Ragged input: a list of 1-D float32 waveforms (CPU or device) -> list of 1-D CPU waveforms, in the input order;
        empty list -> [].  The reference vocodes such a list one utterance at a time (infers/inference_hifigan.py:67).
        Here utterances of similar length (``bucket_padded``: at most ``max_pad`` of a batch is padding) share a batch that
        is zero-padded to its longest member and carries the per-utterance lengths: the front-end reflects every utterance
        at its own end and every generator kernel masks it at its own length, so each result is BIT-IDENTICAL to the
        single-utterance result (tests/test_gpu_parity.py; HiFiGAN and iSTFTNet on the 16-bit path).  The fp32 precision path
        has no ragged kernels and falls back to groups of exactly equal length."""
        outs = [None] * len(wavs)
        h = self.h
        lens = [int(w.shape[-1]) for w in wavs]
        ragged_ok = getattr(self.generator, "precision", None) in (None, "bf16") and all(n > int(h.n_fft) // 2 for n in lens)
        if not ragged_ok:
            for group in bucket_by_length(lens, self.micro_batch):
                batch = torch.stack([wavs[i].reshape(-1).to(torch.float32) for i in group]).to(self.device, non_blocking=True)
                y = self.generator(self.mel(batch)).reshape(len(group), -1).cpu()
                for j, i in enumerate(group):
                    outs[i] = y[j]
            return outs
        up = 1
        for u in h.upsample_rates:
            up *= int(u)
        if getattr(self.generator, "_kind", None) == _lib.GEN_ISTFTNET:
            up *= int(h.gen_istft_hop_size)
        # Host tensors in, host tensors out, two groups in flight: a group's padded batch is assembled in a PINNED staging buffer
        # (plain memcpys), goes up in one asynchronous copy on a side stream and comes back into a pinned buffer on another,
        # so the copies of group g + 1 / g - 1 overlap the kernels of group g (the same choreography as run_host); the results
        # are cut out of the pinned buffer on the host once their copy has landed.  Device-resident inputs skip the staging.
        groups = list(bucket_padded(lens, self.micro_batch, max_pad))
        if not groups:
            return outs
        dev = self.device
        on_host = all(not w.is_cuda for w in wavs)
        if not on_host:
            for group in groups:
                tmax = max(lens[i] for i in group)
                batch = torch.zeros((len(group), tmax), dtype=torch.float32, device=dev)
                for j, i in enumerate(group):
                    batch[j, :lens[i]].copy_(wavs[i].reshape(-1).to(torch.float32), non_blocking=True)
                n = torch.tensor([lens[i] for i in group], dtype=torch.int32, device=dev)
                y = self.vocode(batch, lengths=n).reshape(len(group), -1).cpu()
                for j, i in enumerate(group):
                    outs[i] = y[j, :(1 + lens[i] // int(h.hop_size)) * up].clone()
            return outs
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._stage = None
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        rows = max(len(g) for g in groups)
        t_in = max(max(lens[i] for i in g) for g in groups)
        t_out = (1 + t_in // int(h.hop_size)) * up
        pin = getattr(self, "_list_stage", None)
        if pin is None or pin["rows"] < rows or pin["t_in"] < t_in or pin["t_out"] < t_out:
            pin = self._list_stage = {"rows": rows, "t_in": t_in, "t_out": t_out,
                                      "in": [torch.empty(rows * t_in, dtype=torch.float32).pin_memory() for _ in range(2)],
                                      "out": [torch.empty(rows * t_out, dtype=torch.float32).pin_memory() for _ in range(2)],
                                      "len": [torch.empty(rows, dtype=torch.int32).pin_memory() for _ in range(2)]}
        uploaded, landed, pending = [None, None], [None, None], [None, None]

        def collect(k):  # cut the results of the group that used staging slot k out of its pinned output buffer
            if pending[k] is None:
                return
            group, hout = pending[k]
            landed[k].synchronize()
            for j, i in enumerate(group):
                outs[i] = hout[j, :(1 + lens[i] // int(h.hop_size)) * up].clone()
            pending[k] = None

        for g, group in enumerate(groups):
            k = g & 1
            collect(k)                      # slot k's previous results are read before its buffers are reused
            if uploaded[k] is not None:
                uploaded[k].synchronize()   # ... and its previous input has left the pinned buffer
            nb, tmax = len(group), max(lens[i] for i in group)
            hin = pin["in"][k][:nb * tmax].view(nb, tmax)
            hlen = pin["len"][k][:nb]
            for j, i in enumerate(group):
                hin[j, :lens[i]].copy_(wavs[i].reshape(-1))
                hin[j, lens[i]:].zero_()
                hlen[j] = lens[i]
            with torch.cuda.stream(self._h2d):
                batch = hin.to(dev, non_blocking=True)
                n = hlen.to(dev, non_blocking=True)
                uploaded[k] = torch.cuda.Event()
                uploaded[k].record(self._h2d)
            cur.wait_event(uploaded[k])
            y = self.vocode(batch, lengths=n).reshape(nb, -1)
            batch.record_stream(cur)
            n.record_stream(cur)
            done = torch.cuda.Event()
            done.record(cur)
            hout = pin["out"][k][:nb * y.shape[1]].view(nb, y.shape[1])
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done)
                hout.copy_(y, non_blocking=True)
                y.record_stream(self._d2h)
                landed[k] = torch.cuda.Event()
                landed[k].record(self._d2h)
            pending[k] = (group, hout)
        collect(0)
        collect(1)
        cur.wait_stream(self._d2h)
        return outs
