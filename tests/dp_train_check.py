#!/usr/bin/env python
"""Data-parallel training check on N GPUs of one box (NCCL): every rank runs the generator + mel-L1 step on its own
half of a batch through the CUDA forward/backward, gradients are averaged with ONE flat all-reduce
(nvse.allreduce_gradients), and the result is compared with a single-GPU step over the whole batch.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dp_train_check.py"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import build_generator, pkg  # noqa: E402
import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.HIFIGAN_V1
a = cfg
margs = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["sampling_rate"] / 2)
state = synth.make_state(cfg, 1234, "init")
per = 8                                   # segments per rank (global batch = per * world)
mel_all = torch.from_numpy(synth.make_mel(per * world, 32, 1)).to(dev)
y_all = torch.from_numpy(synth.make_wave(per * world, 8192, 2)).to(dev)


def step(gen, sl):
    gen.zero_grad(set_to_none=True)
    loss = F.l1_loss(pkg.mel_spectrogram(y_all[sl], *margs), pkg.mel_spectrogram(gen(mel_all[sl]), *margs)) * 45
    loss.backward()
    return loss.detach()


gen = build_generator(cfg, state, dev).train()
gen.precision = "fp32"
mine = slice(rank * per, (rank + 1) * per)
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    loss = step(gen, mine)
    e1.record()
    n = pkg.allreduce_gradients(list(gen.parameters()))
    e2.record()
    torch.cuda.synchronize()
dp = {k: p.grad.clone() for k, p in gen.named_parameters()}
ms_step, ms_ar = e0.elapsed_time(e1), e1.elapsed_time(e2)

# single-GPU reference over the whole batch (mean-reduced L1 over equal shards == mean of the per-shard losses)
ref_gen = build_generator(cfg, state, dev).train()
ref_gen.precision = "fp32"
step(ref_gen, slice(0, per * world))
worst = 0.0
for k, p in ref_gen.named_parameters():
    worst = max(worst, float((dp[k] - p.grad).abs().max() / (p.grad.abs().max() + 1e-20)))
w = torch.tensor([worst], device=dev)
dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_train_check: world {world}, {n} gradient elements ({n * 4 / 1e6:.1f} MB) in one all-reduce: {ms_ar:.3f} ms "
          f"(step {ms_step:.2f} ms); worst per-tensor max deviation from the single-GPU whole-batch gradient {float(w):.2e}")
    assert float(w) <= 1e-4, float(w)
dist.destroy_process_group()
