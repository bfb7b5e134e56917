#!/usr/bin/env python
"""Per-kernel device time of one batch-1 forward (2 s utterance) through the library's event profiler."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator, pkg, lib_mod  # noqa: E402
import synth  # noqa: E402
sec = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda:0", remove_wn=True)
wav = torch.from_numpy(synth.make_wave(B, int(sec * 22050), 0)).to("cuda:0")
mel = lambda y: pkg.mel_spectrogram(y, cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["fmax"])
with torch.no_grad():
    for _ in range(3): gen(mel(wav))
    torch.cuda.synchronize()
    lib_mod.profile_begin()
    for _ in range(10): gen(mel(wav))
    torch.cuda.synchronize()
prof = lib_mod.profile_end()
tot = 0.0
for p in sorted(prof, key=lambda r: -r["ms"]):
    print(f"{p['kernel']:24s} x{p['launches'] // 10:3d}  {p['ms'] / 10 * 1e3:8.1f} us per forward")
    tot += p["ms"] / 10
print(f"sum of kernel times {tot * 1e3:.1f} us per forward")
