import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import torch, numpy as np
import __graft_entry__ as e
pkg = e.load_package()
import synth
from util import build_generator
for cfg, F in ((synth.HIFIGAN_V1, 37), (synth.ISTFTNET, 21)):
    state = synth.make_state(cfg, 1234, "init")
    gen = build_generator(cfg, state, "cuda:0", remove_wn=True)
    mel = torch.from_numpy(synth.make_mel(2, F, 3)).to("cuda:0")
    for prec in ("bf16", "fp32"):
        gen.precision = prec
        with torch.no_grad():
            y = gen(mel)
        torch.cuda.synchronize()
        print(cfg["model_name"], prec, tuple(y.shape), float(y.abs().max()))
print("abort flag:", pkg._lib.tc_abort_status())
