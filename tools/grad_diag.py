#!/usr/bin/env python
"""Per-tensor difference between the tensor-core ('bf16') and the fp32 training path (tests' hifigan_train fixture)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import build_generator  # noqa: E402
import synth  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "grads_hifigan_train_f6"
gold = synth.load_golden(name)
meta = gold["meta"]
cfg = synth.CONFIGS[meta["cfg"]]
state = synth.make_state(cfg, meta["weight_seed"], meta["regime"])
res = {}
for prec in ("fp32", "bf16"):
    gen = build_generator(cfg, state, "cuda").train()
    gen.precision = prec
    x = torch.from_numpy(gold["mel"]).cuda().requires_grad_(True)
    (gen(x) * torch.from_numpy(gold["dout"]).cuda()).sum().backward()
    res[prec] = {k: p.grad.clone() for k, p in gen.named_parameters()}
rows = []
for k, g in res["fp32"].items():
    h = res["bf16"][k]
    rows.append((float((h - g).norm() / (g.norm() + 1e-20)), float(torch.nn.functional.cosine_similarity(h.flatten(), g.flatten(), dim=0)), k, float(g.norm())))
rows.sort(reverse=True)
for r in rows[:12]:
    print(f"rel L2 {r[0]:.3e}  cos {r[1]:.6f}  |g| {r[3]:.3e}  {r[2]}")
print("median rel L2", sorted(r[0] for r in rows)[len(rows) // 2])
