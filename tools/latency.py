#!/usr/bin/env python
"""Single-utterance latency (the reference script's batch-1 loop, BASELINE cfg1: 2 s -> F = 173) through the drop-in
modules: wav -> mel_spectrogram -> generator, per call, CUDA-event timed.  usage: latency.py [seconds=2]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator, pkg  # noqa: E402
import synth  # noqa: E402
sec = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda:0", remove_wn=True)
wav = torch.from_numpy(synth.make_wave(1, int(sec * 22050), 0)).to("cuda:0")
mel = lambda y: pkg.mel_spectrogram(y, cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["fmax"])
for prec in ("bf16", "fp32"):
    gen.precision = prec
    with torch.no_grad():
        for _ in range(5): y = gen(mel(wav))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(50): y = gen(mel(wav))
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / 50
    print(f"{prec}: {sec:g} s utterance, batch 1: {ms:.3f} ms per call (device), {(t1 - t0) * 1e3 / 50:.3f} ms wall -> {y.shape[-1] / 22050 / (ms * 1e-3):.0f} audio-s/s")
