#!/usr/bin/env python
"""Vocoder.run_list on a ragged list (uniform 5 .. 10 s utterances, host tensors in, host tensors out) next to run_host on the
same number of audio seconds at one fixed length: what padding, bucketing and the per-group host work cost."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg, build_generator  # noqa: E402
import synth  # noqa: E402
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda", remove_wn=True)
gen.precision = "bf16"
voc = pkg.Vocoder(gen, synth.AttrDict(cfg), micro_batch=32, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
lens = rng.integers(5 * 22050, 10 * 22050, size=n)
wavs = [torch.from_numpy(synth.make_wave(1, int(t), 100 + i)[0]).pin_memory() for i, t in enumerate(lens)]
audio_s = float(sum((1 + int(t) // 256) * 256 for t in lens)) / 22050
for _ in range(2):
    outs = voc.run_list(wavs)
torch.cuda.synchronize()
t0 = time.perf_counter()
outs = voc.run_list(wavs)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"run_list: {n} utterances of 5 .. 10 s ({audio_s:.0f} audio-s): {dt * 1e3:.1f} ms = {audio_s / dt:.0f} audio-s/s")
T = int(np.mean(lens))
fixed = torch.from_numpy(synth.make_wave(n, T, 7)).pin_memory()
out = None
for _ in range(2):
    out = voc.run_host(fixed, out)
torch.cuda.synchronize()
t0 = time.perf_counter()
voc.run_host(fixed, out)
torch.cuda.synchronize()
dt2 = time.perf_counter() - t0
a2 = n * (1 + T // 256) * 256 / 22050
print(f"run_host: {n} x {T / 22050:.2f} s ({a2:.0f} audio-s): {dt2 * 1e3:.1f} ms = {a2 / dt2:.0f} audio-s/s")
for T2 in (T // 4 * 4, 220500):
    fx = torch.from_numpy(synth.make_wave(128, T2, 7)).pin_memory()
    o2 = None
    for _ in range(2):
        o2 = voc.run_host(fx, o2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        voc.run_host(fx, o2)
    torch.cuda.synchronize()
    d = (time.perf_counter() - t0) / 3
    a3 = 128 * (1 + T2 // 256) * 256 / 22050
    print(f"run_host: 128 x {T2} samples ({a3:.0f} audio-s): {d * 1e3:.1f} ms = {a3 / d:.0f} audio-s/s")
