#!/usr/bin/env python
"""Discriminator half of the reference's training step (train_time_wi_inv.py:188-236) at its shape (batch 16 x 8192 samples):
D step (forward on y / y_hat.detach(), ls loss, backward) + G-step part (forward, feature + ls generator loss, backward to y_hat),
this repo's kernels vs the same module tree on stock PyTorch (cuDNN) with TF32 on (PyTorch's default for convolutions) and off."""
import os, sys, json
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import pkg, lib_mod  # noqa: E402
models = pkg.Models.models
dev = "cuda:0"
B, T = int(os.environ.get("DB_B", 16)), int(os.environ.get("DB_T", 8192))
torch.manual_seed(0)
y = (torch.rand(B, T, device=dev) - 0.5)
yh = (torch.rand(B, T, device=dev) - 0.5).requires_grad_(True)
torch.manual_seed(1)
mpd = models.MultiPeriodDiscriminator([2, 3, 5, 7, 11]).to(dev).train()
msd = models.MultiScaleDiscriminator().to(dev).train()


def step():
    for net in (mpd, msd):
        d_r, d_g, _, _ = net(y, yh.detach())
        models.ls_discriminator_loss(d_r, d_g)[0].backward()
    mpd.zero_grad(set_to_none=True); msd.zero_grad(set_to_none=True)
    loss = 0
    for net in (mpd, msd):
        d_r, d_g, f_r, f_g = net(y, yh)
        loss = loss + models.ls_generator_loss(d_g)[0] + models.feature_loss(f_r, f_g)
    loss.backward()
    yh.grad = None
    mpd.zero_grad(set_to_none=True); msd.zero_grad(set_to_none=True)


def timed(n=5):
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"shape": [B, T]}
lib_mod.profile_begin()
step()
prof = lib_mod.profile_end()
res["ours_ms"] = timed()
res["kernels"] = sorted(prof, key=lambda r: -r["ms"])[:12]
saved = (models._DiscConv1d.forward, models._DiscConv2d.forward, models._MeanPool.forward)
lrelu = lambda x, s: x if s == 1.0 else torch.nn.functional.leaky_relu(x, s)
models._DiscConv1d.forward = lambda self, x, slope=1.0: lrelu(torch.nn.Conv1d.forward(self, x), slope)
models._DiscConv2d.forward = lambda self, x, slope=1.0: lrelu(torch.nn.Conv2d.forward(self, x), slope)
models._MeanPool.forward = lambda self, x: torch.nn.functional.avg_pool1d(x, self.kernel_size, self.stride, self.padding)
torch.backends.cudnn.allow_tf32 = True
res["stock_tf32_ms"] = timed()
torch.backends.cudnn.allow_tf32 = False
res["stock_fp32_ms"] = timed()
print(json.dumps(res, indent=1))
