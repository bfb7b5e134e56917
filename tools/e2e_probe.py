#!/usr/bin/env python
"""Where the host-to-host leg of bench.py spends its time: PCIe copy bandwidths alone, run_device, run_host."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg  # noqa: E402
import synth  # noqa: E402
dev = "cuda:0"
U, T = 128, 220160
cfg = synth.HIFIGAN_V1
h = synth.AttrDict(cfg)
gen = pkg.HiFiGAN(h).to(dev).eval(); gen.remove_weight_norm()
voc = pkg.Vocoder(gen, h, micro_batch=32, device=dev)
wav_host = torch.from_numpy(synth.make_wave(U, T, 1)).pin_memory()
wav_dev = wav_host.to(dev)
out_dev = voc.run_device(wav_dev)
out_host = torch.empty(out_dev.shape, dtype=torch.float32).pin_memory()
def timed(fn, n=5):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
tmp = torch.empty_like(wav_dev)
for name, fn in (("h2d 113 MB", lambda: tmp.copy_(wav_host, non_blocking=True)), ("d2h 113 MB", lambda: out_host.copy_(out_dev, non_blocking=True)),
                 ("run_device", lambda: voc.run_device(wav_dev, out_dev)), ("run_host", lambda: voc.run_host(wav_host, out_host)),
                 ("run_device again", lambda: voc.run_device(wav_dev, out_dev)), ("run_host again", lambda: voc.run_host(wav_host, out_host))):
    fn(); fn()
    ev, wall = timed(fn)
    print(f"{name:18s} {ev:8.2f} ms (events)  {wall:8.2f} ms (wall)")
