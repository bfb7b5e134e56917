#!/usr/bin/env python
"""Repeat-run stress of the fused kernels: N forwards at several batch sizes, every output compared bit for bit
with the first of its shape; reports the abort flag.  usage: stress.py [iters=200]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator, lib_mod  # noqa: E402
import synth  # noqa: E402
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bad = 0
for cfg in (synth.HIFIGAN_V1, synth.ISTFTNET):
    gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda:0", remove_wn=True)
    gen.precision = "bf16"
    for B, F in ((1, 173), (5, 431), (32, 862)):
        mel = torch.from_numpy(synth.make_mel(B, F, 7)).to("cuda:0")
        with torch.no_grad():
            first = gen(mel).clone()
            n = max(3, iters // (B * F // 173 + 1))
            for i in range(n):
                if not torch.equal(gen(mel), first):
                    bad += 1
        torch.cuda.synchronize()
        print(cfg["model_name"], B, F, "runs", n, "mismatches so far", bad, "abort", lib_mod.tc_abort_status(reset=False), flush=True)
print("OK" if bad == 0 and not lib_mod.tc_abort_status() else "FAILED")
