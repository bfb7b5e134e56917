#!/usr/bin/env python
"""Error map of the fused ResBlock kernel vs a torch reference rounding at the same points."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import lib_mod, resblock1_cl
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
C_, k, T, B = (int(v) for v in sys.argv[1:5])
dils = [int(v) for v in sys.argv[5].split(",")]
def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)
_bf = lambda x: x.to(torch.bfloat16).to(torch.float32)
n = len(dils)
x = _rand((B, C_, T), 41)
w1 = [_rand((C_, C_, k), 42 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
w2 = [_rand((C_, C_, k), 52 + m, 1.0 / np.sqrt(C_ * k)) for m in range(n)]
b1 = [_rand((C_,), 62 + m, 0.3) for m in range(n)]
b2 = [_rand((C_,), 72 + m, 0.3) for m in range(n)]
def reference(dev, rnd):
    r = x.to(dev)
    f = _bf if rnd else (lambda v: v)
    for m, d in enumerate(dils):
        h = F.conv1d(f(F.leaky_relu(r, 0.1)), f(w1[m].to(dev)), b1[m].to(dev), dilation=d, padding=(k - 1) * d // 2)
        r = F.conv1d(f(F.leaky_relu(h, 0.1)), f(w2[m].to(dev)), b2[m].to(dev), padding=(k - 1) // 2) + r
    return r.to(DEV)
ref = reference(DEV, True)
ref_cpu = reference("cpu", True)
ref32 = reference(DEV, False)
out = resblock1_cl(x, w1, b1, w2, b2, dils)
out_b = resblock1_cl(x, w1, b1, w2, b2, dils)
print("deterministic:", bool(torch.equal(out, out_b)), " aborted:", lib_mod.tc_abort_status())
print(f"|ref| mean {float(ref.abs().mean()):.3f}; gpu-ref vs cpu-ref mean {float((ref-ref_cpu).abs().mean()):.2e}; "
      f"bf16-ref vs fp32-ref mean {float((ref-ref32).abs().mean()):.2e}")
for name, r in (("gpu-ref", ref), ("cpu-ref", ref_cpu), ("fp32-ref", ref32)):
    e = (out - r).abs()
    print(f"kernel vs {name}: max {float(e.max()):.2e} mean {float(e.mean()):.2e}")
e = (out - ref_cpu).abs()
seg = 64
for b in range(B):
    row = [f"{float(e[b, :, t:t + seg].mean()):.1e}" for t in range(0, T, seg)]
    print(f"b={b} mean err per {seg} steps:", " ".join(row))
