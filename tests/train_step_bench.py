#!/usr/bin/env python
"""Generator part of one training step at the reference's training shape (cfgs/hifigan_v1_config.json: batch 16,
segment 8192 -> 32 frames; train_time_wi_inv.py:166-179,222-236): y_g = G(mel); L = 45 * L1(mel(y), mel(y_g));
L.backward(); AdamW step -- through the CUDA forward/backward of this repo, and through stock PyTorch (the oracle port, i.e. what
the reference's nn.Module dispatches: cuDNN + element-wise kernels) on the same GPU.  CUDA-event timed.
usage: train_bench.py [batch=16] [frames=32] [steps=10]"""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import build_generator, pkg, lib_mod  # noqa: E402
import synth  # noqa: E402
from oracle import torch_port  # noqa: E402  (checker / comparison arm only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
FR = int(sys.argv[2]) if len(sys.argv) > 2 else 32
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = synth.HIFIGAN_V1
a = cfg
margs = (a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["sampling_rate"] / 2)
state = synth.make_state(cfg, 1234, "init")
mel_in = torch.from_numpy(synth.make_mel(B, FR, 1)).cuda()
y = torch.from_numpy(synth.make_wave(B, FR * 256, 2)).cuda()


HOST_MS = {}


def timed(fn, steps, tag=None):
    import time
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    t1 = time.perf_counter()   # host time to ENQUEUE the steps (no sync inside): >= device time means host-bound
    torch.cuda.synchronize()
    if tag:
        HOST_MS[tag] = (t1 - t0) * 1e3 / steps
    return e0.elapsed_time(e1) / steps


gen = build_generator(cfg, state, "cuda").train()
y_mel = pkg.mel_spectrogram(y, *margs)
LR = 1e-7  # the optimiser step is part of the step (train_time_wi_inv.py:236); tiny so that the timed steps stay comparable
opt_ours = torch.optim.AdamW(gen.parameters(), LR, betas=(0.8, 0.99))


def ours():
    opt_ours.zero_grad(set_to_none=True)
    y_g = gen(mel_in)
    loss = F.l1_loss(y_mel, pkg.mel_spectrogram(y_g, *margs)) * 45
    loss.backward()
    opt_ours.step()
    return loss


leaves = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in state.items()}
torch_port._basis_cache.clear()


def mel_torch(w):
    key = "gpu"
    if key not in torch_port._basis_cache:
        basis = torch.from_numpy(torch_port.np_oracle.mel_filterbank(a["sampling_rate"], a["n_fft"], a["num_mels"], a["fmin"], margs[-1])).cuda()
        torch_port._basis_cache[key] = (basis, torch.hann_window(a["win_size"], device="cuda"))
    basis, window = torch_port._basis_cache[key]
    spec = torch.stft(w, a["n_fft"], hop_length=a["hop_size"], win_length=a["win_size"], window=window, center=True, return_complex=True)
    return torch.log(torch.clamp(basis @ spec.abs(), min=1e-5))


y_mel_t = mel_torch(y)


opt_stock = torch.optim.AdamW(list(leaves.values()), LR, betas=(0.8, 0.99))


def stock():
    opt_stock.zero_grad(set_to_none=True)
    y_g = torch_port.hifigan_forward_autograd(torch_port.fold_state(leaves), cfg, mel_in)
    loss = F.l1_loss(y_mel_t, mel_torch(y_g)) * 45
    loss.backward()
    opt_stock.step()
    return loss


res = {"batch": B, "frames": FR, "segment": FR * 256}
l_ours, l_stock = float(ours().detach()), float(stock().detach())
res["loss_ours"], res["loss_stock_torch"] = l_ours, l_stock
for prec in ("fp32", "bf16"):
    gen.precision = prec
    res[f"ours_{prec}_ms"] = timed(ours, STEPS, f"ours_{prec}")
    res[f"loss_ours_{prec}"] = float(ours().detach())
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    res["stock_torch_ms_tf32" if tf32 else "stock_torch_ms_fp32"] = timed(stock, STEPS, "stock_tf32" if tf32 else "stock_fp32")
res["host_enqueue_ms_per_step"] = HOST_MS
# where the time goes (per-launch CUDA events of the library's own profiler)
lib_mod.profile_begin()
ours()
torch.cuda.synchronize()
prof = lib_mod.profile_end()
res["profile"] = sorted(prof, key=lambda r: -r["ms"])[:14]
print(json.dumps(res, indent=1))
