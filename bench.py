#!/usr/bin/env python
"""Headline benchmark: HiFi-GAN V1 audio-seconds per second on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                      (the reference's CPU path, oracle port)

One step = one pass of the hot path (wav -> mel_spectrogram -> HiFi-GAN V1 -> wav) over this
rank's shard of synthetic utterances: per GPU, 128 utterances x 10 s (cfg5 of BASELINE.json is
1024 such utterances over 8 GPUs, so N = 8 is exactly cfg5 and per-GPU work is fixed: weak
scaling, no collective on the data path), processed in micro-batches of 32.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

SR = 22050
FLOP_PER_FRAME = 2 * 307_052_544  # HiFi-GAN V1, SURVEY.md 8(d) / BASELINE.md 2
METRIC = "HiFi-GAN V1 audio-sec/sec (wav -> mel -> generator -> wav)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--utts-per-gpu", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--cpu-utts", type=int, default=2, help="utterances in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--total-utts", type=int, default=0,
                    help="strong scaling: a FIXED pool of this many utterances (cfg5: 1024) split with shard_range over the ranks; "
                         "0 = weak scaling (every rank its own --utts-per-gpu shard)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (cfg1 / cfg2 / eager competitor / train step)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


REF_STAGE = os.path.join(ROOT, "baseline", "_ref")


def load_reference_modules():
    """The reference's OWN dataset.py / Models/hifigan.py, loaded by file path from baseline/_ref (a plain copy of the
    reference files staged by __graft_entry__.build(); the reference has no setup.py to pip-install).  librosa is not in
    this image: tests/shims supplies librosa.filters.mel (= oracle.np_oracle.mel_filterbank).  None when not staged."""
    import importlib.util
    if not os.path.isfile(os.path.join(REF_STAGE, "Models", "hifigan.py")):
        return None
    shims = os.path.join(ROOT, "tests", "shims")
    if shims not in sys.path:
        sys.path.append(shims)
    mods = {}
    for name, rel in (("_bench_ref_dataset", "dataset.py"), ("_bench_ref_hifigan", os.path.join("Models", "hifigan.py"))):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_STAGE, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["_bench_ref_dataset"], mods["_bench_ref_hifigan"]


def reference_generator(hifigan_mod, cfg, device="cpu"):
    """Models.HiFiGAN(h) of the reference with the bench's synthetic state dict, eval(), weight norm removed."""
    import contextlib
    import synth
    gen = hifigan_mod.HiFiGAN(synth.AttrDict(cfg))
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state(cfg, 1234, "init").items()})
    gen = gen.to(device).eval()
    with contextlib.redirect_stdout(sys.stderr):
        gen.remove_weight_norm()
    return gen


def cpu_reference(args, n_utts, steps, warmup, seconds=None):
    """The reference's CPU path, one utterance at a time like infers/inference_hifigan.py:67-84, fp32, all host threads:
    the reference's own modules when baseline/_ref is staged (kind "reference"), else the oracle port (same torch.stft /
    oneDNN conv calls; kind "port")."""
    import synth
    cfg = synth.HIFIGAN_V1
    torch.set_num_threads(os.cpu_count() or 1)
    T = int((seconds or args.seconds) * SR)
    wav = torch.from_numpy(synth.make_wave(n_utts, T, 0))
    margs = (cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["fmax"])
    ref = load_reference_modules()
    if ref is not None:
        ds, hg = ref
        gen = reference_generator(hg, cfg)
        kind = "reference"

        @torch.no_grad()
        def step():
            n = 0
            for u in range(n_utts):
                n += gen(ds.mel_spectrogram(wav[u:u + 1], *margs)).shape[-1]
            return n
    else:
        from oracle import torch_port
        folded = torch_port.fold_state(synth.make_state(cfg, 1234, "init"))
        kind = "port"

        def step():
            n = 0
            for u in range(n_utts):
                n += torch_port.hifigan_forward(folded, cfg, torch_port.mel_spectrogram(wav[u:u + 1], *margs)).shape[-1]
            return n

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    samples = 0
    for _ in range(steps):
        samples += step()
    dt = time.perf_counter() - t0
    return samples / SR / dt, dt / steps, torch.get_num_threads(), kind


def train_step(pkg, synth, cfg, dev, batch=16, frames=32, steps=5):
    """Secondary number (not the headline metric): the generator part of the reference's training
    step at its own shape (cfgs/hifigan_v1_config.json: batch 16 x segment 8192) through the CUDA forward/backward:
    y_g = G(mel); L = 45 * L1(mel(y), mel(y_g)); L.backward()  (train_time_wi_inv.py:166-179,222-236)."""
    import torch.nn.functional as F
    gen = pkg.HiFiGAN(synth.AttrDict(cfg))
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state(cfg, 1234, "init").items()})
    gen = gen.to(dev).train()
    margs = (cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["sampling_rate"] / 2)
    mel_in = torch.from_numpy(synth.make_mel(batch, frames, 1)).to(dev)
    y_mel = pkg.mel_spectrogram(torch.from_numpy(synth.make_wave(batch, frames * 256, 2)).to(dev), *margs)

    opt = torch.optim.AdamW(gen.parameters(), 1e-7, betas=(0.8, 0.99))  # the weights change every step, as in training

    def step():
        opt.zero_grad(set_to_none=True)
        (F.l1_loss(y_mel, pkg.mel_spectrogram(gen(mel_in), *margs)) * 45).backward()
        opt.step()

    out = {"config": f"HiFi-GAN V1 generator + mel-L1, batch {batch} x {frames * 256} samples, forward + backward + AdamW step + per-step weight upload"}
    for prec in ("bf16", "fp32"):
        gen.precision = prec
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pkg._lib.launch_count()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[f"ms_{prec}"] = e0.elapsed_time(e1) / steps
        out[f"gpu_launches_{prec}"] = int((pkg._lib.launch_count() - l0) / steps)
    out["note"] = "bf16 = MRF convs, dgrad and wgrad on tcgen05 (bf16 operands, fp32 accumulate); fp32 = CUDA cores, the parity path"
    return out


def disc_step(pkg, dev, batch=16, samples=8192, steps=3):
    """Secondary number: the discriminator half of the reference's training step at its own shape
    (train_time_wi_inv.py:188-236: D step on (y, y_g.detach()) + the generator step's pass through MPD and MSD with the
    feature and least-squares losses, backward to y_g), this repo's fp32 kernels (csrc/disc.cu) next to the SAME module
    tree on stock PyTorch (cuDNN) with TF32 on (PyTorch's default, what the reference trains with) and off."""
    models = pkg.Models.models
    torch.manual_seed(0)
    y = torch.rand(batch, samples, device=dev) - 0.5
    yh = (torch.rand(batch, samples, device=dev) - 0.5).requires_grad_(True)
    torch.manual_seed(1)
    nets = [models.MultiPeriodDiscriminator([2, 3, 5, 7, 11]).to(dev).train(), models.MultiScaleDiscriminator().to(dev).train()]

    def step():
        for net in nets:
            d_r, d_g, _, _ = net(y, yh.detach())
            models.ls_discriminator_loss(d_r, d_g)[0].backward()
            net.zero_grad(set_to_none=True)
        loss = 0
        for net in nets:
            d_r, d_g, f_r, f_g = net(y, yh)
            loss = loss + models.ls_generator_loss(d_g)[0] + models.feature_loss(f_r, f_g)
        loss.backward()
        yh.grad = None
        for net in nets:
            net.zero_grad(set_to_none=True)

    def timed():
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pkg._lib.launch_count()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, int((pkg._lib.launch_count() - l0) / steps)

    out = {"config": f"MPD (periods 2,3,5,7,11) + MSD, batch {batch} x {samples} samples: D step + G-step pass, forward + backward"}
    out["ms_fp32"], out["gpu_launches"] = timed()
    saved = (models._DiscConv1d.forward, models._DiscConv2d.forward, models._MeanPool.forward, torch.backends.cudnn.allow_tf32)
    lrelu = lambda x, sl: x if sl == 1.0 else torch.nn.functional.leaky_relu(x, sl)  # noqa: E731
    try:
        models._DiscConv1d.forward = lambda self, x, slope=1.0: lrelu(torch.nn.Conv1d.forward(self, x), slope)
        models._DiscConv2d.forward = lambda self, x, slope=1.0: lrelu(torch.nn.Conv2d.forward(self, x), slope)
        models._MeanPool.forward = lambda self, x: torch.nn.functional.avg_pool1d(x, self.kernel_size, self.stride, self.padding)
        torch.backends.cudnn.allow_tf32 = True
        out["stock_pytorch_tf32_ms"] = timed()[0]
        torch.backends.cudnn.allow_tf32 = False
        out["stock_pytorch_fp32_ms"] = timed()[0]
    finally:
        models._DiscConv1d.forward, models._DiscConv2d.forward, models._MeanPool.forward, torch.backends.cudnn.allow_tf32 = saved
    out["note"] = "fp32 CUDA-core kernels (bit-reproducible); cuDNN's TF32 tensor-core path is faster, see DESIGN.md"
    return out


def cfg4_istftnet(pkg, synth, dev, batch=32, frames=690, reps=5):
    """cfg4 of BASELINE.json: the iSTFTNet generator (Models/istftnet.py:271-328) on mel [32, 80, 690] (32 x 8 s), 16-bit
    tensor-core path, device-resident; algorithmic FLOPs = 4.116e8 per frame."""
    cfg = synth.ISTFTNET
    gen = pkg.iSTFTNet(synth.AttrDict(cfg))
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state(cfg, 1234, "init").items()})
    gen = gen.to(dev).eval()
    gen.remove_weight_norm()
    gen.precision = "bf16"
    mel = torch.from_numpy(synth.make_mel(batch, frames, 1)).to(dev)
    with torch.no_grad():
        for _ in range(3):
            out = gen(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gen(mel)
        e1.record()
        torch.cuda.synchronize()
        gen.precision = "fp32"
        ref = gen(mel[:2]).double()
    ms = e0.elapsed_time(e1) / reps
    deg = out[:2].double()
    ref, deg = ref - ref.mean(-1, keepdim=True), deg - deg.mean(-1, keepdim=True)
    snr = float((10 * torch.log10(ref.pow(2).sum(-1) / (ref - deg).pow(2).sum(-1).clamp_min(1e-30))).min())
    return {"config": f"cfg4: iSTFTNet, mel [{batch}, 80, {frames}] -> wav, device-resident, 16-bit tensor-core path", "ms": ms,
            "value": batch * out.shape[-1] / SR / (ms * 1e-3), "unit": "audio-sec/sec", "tflops_algorithmic": 4.116e8 * batch * frames / ms / 1e9,
            "parity_snr_db_vs_fp32_path": snr}


def cfg3_hifigan(gen, synth, dev, pk, precision, batch=32, frames=690, reps=5):
    """cfg3 of BASELINE.json: the HiFi-GAN V1 generator alone on mel [32, 80, 690] (32 x 8 s), device-resident; BASELINE.md 2:
    1.3559e13 FLOP, the north star's >= 50 % of the sustained bf16 peak means <= 19.4 ms."""
    mel = torch.from_numpy(synth.make_mel(batch, frames, 1)).to(dev)
    with torch.no_grad():
        for _ in range(3):
            gen(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = gen(mel)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = FLOP_PER_FRAME * batch * frames / ms / 1e9
    return {"config": f"cfg3: HiFi-GAN V1 generator, mel [{batch}, 80, {frames}] -> wav, device-resident, precision {precision}", "ms": ms,
            "value": batch * out.shape[-1] / SR / (ms * 1e-3), "unit": "audio-sec/sec", "tflops_algorithmic": tf,
            "frac_of_sustained_bf16_peak": tf / pk["bf16_tflops_sustained"]}


def cfg1_gpu(voc, synth, dev, reps=30):
    """cfg1 of BASELINE.json on the GPU through the public host-to-host call: batch 1, 2 s -> mel -> HiFi-GAN V1 -> wav."""
    wav = torch.from_numpy(synth.make_wave(1, 2 * SR, 3)).pin_memory()
    out = None
    for _ in range(5):
        out = voc.run_host(wav, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        voc.run_host(wav, out)
        torch.cuda.current_stream(dev).synchronize()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"config": "cfg1: HiFi-GAN V1, batch 1, 2 s (173 frames -> 44288 samples), pinned host wav -> pinned host wav, wall clock incl. sync",
            "ms": med * 1e3, "value": out.shape[-1] / SR / med, "unit": "audio-sec/sec"}


def cfg2_frontend(voc, synth, cfg, dev, pk):
    """cfg2 of BASELINE.json: mel_spectrogram alone on 64 x 4 s, warm (working set L2-resident, back to back) and cold (a
    256 MB write between launches evicts L2), plus a ~1 GB steady-state size; algorithmic bytes = 4BT + 4*B*80*F."""
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, B, secs, reps in (("cfg2_64x4s", 64, 4.0, 50), ("steady_2048x4s", 2048, 4.0, 10)):
        T = int(secs * SR)
        wav = torch.from_numpy(synth.make_wave(8, T, 9)).to(dev).repeat(B // 8, 1).contiguous()
        nbytes = 4.0 * B * T + 4.0 * B * 80 * (1 + T // cfg["hop_size"])
        for _ in range(3):
            voc.mel(wav)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            voc.mel(wav)
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / reps
        cold = []
        for _ in range(min(reps, 20)):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            voc.mel(wav)
            b.record()
            torch.cuda.synchronize()
            cold.append(a.elapsed_time(b))
        cold_ms = statistics.median(cold)
        out[name] = {"bytes": nbytes, "warm_us": warm * 1e3, "warm_gbs": nbytes / warm / 1e6, "warm_frac_hbm": nbytes / warm / 1e6 / pk["hbm_gbs"],
                     "cold_us": cold_ms * 1e3, "cold_gbs": nbytes / cold_ms / 1e6, "cold_frac_hbm": nbytes / cold_ms / 1e6 / pk["hbm_gbs"]}
        del wav
    out["note"] = "CUDA events around each launch; cold = L2 flushed by a 256 MB fill before every launch; peak = measured copy bandwidth"
    return out


def eager_competitor(args, synth, cfg, dev, wav_dev):
    """BASELINE.md 3's on-box competitor: the reference's modules under stock PyTorch eager on the SAME B200 (cuFFT /
    cuDNN / cuBLAS), one micro-batch of the workload, fp32 as shipped (PyTorch's default lets cuDNN use TF32) and under
    torch.autocast(bf16)."""
    ref = load_reference_modules()
    n = min(args.micro_batch, wav_dev.shape[0])
    wav = wav_dev[:n]
    margs = (cfg["n_fft"], cfg["num_mels"], cfg["sampling_rate"], cfg["hop_size"], cfg["win_size"], cfg["fmin"], cfg["fmax"])
    if ref is not None:
        ds, hg = ref
        gen = reference_generator(hg, cfg, dev)
        fwd, kind = (lambda: gen(ds.mel_spectrogram(wav, *margs))), "reference modules (baseline/_ref)"
    else:
        from oracle import torch_port
        folded = {k: v.to(dev) for k, v in torch_port.fold_state(synth.make_state(cfg, 1234, "init")).items()}
        torch_port._basis_cache.clear()

        def fwd():
            basis = torch.from_numpy(__import__("oracle.np_oracle", fromlist=["x"]).mel_filterbank(cfg["sampling_rate"], cfg["n_fft"], cfg["num_mels"], cfg["fmin"], cfg["fmax"])).to(dev)
            spec = torch.stft(wav, cfg["n_fft"], hop_length=cfg["hop_size"], win_length=cfg["win_size"], window=torch.hann_window(cfg["win_size"], device=dev),
                              center=True, return_complex=True)
            return torch_port.hifigan_forward(folded, cfg, torch.log(torch.clamp(basis @ spec.abs(), min=1e-5)))
        kind = "oracle port"
    res = {"impl": kind, "sample": f"{n} x {args.seconds:g} s utterances per step (one micro-batch), wav on the device -> wav on the device"}
    for label, ctx in (("fp32_tf32_default", None), ("autocast_bf16", torch.bfloat16)):
        def run():
            with torch.no_grad():
                if ctx is None:
                    return fwd()
                with torch.autocast("cuda", dtype=ctx):
                    return fwd()
        for _ in range(2):
            y = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[label] = {"ms": ms, "value": n * y.shape[-1] / SR / (ms * 1e-3), "unit": "audio-sec/sec"}
        del y
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nranks = max(world, args.gpus)
    if args.total_utts > 0:  # strong scaling: the fixed cfg5 pool split per utterance (shard.shard_range)
        workload = (f"cfg5 pool: {args.total_utts} x {args.seconds:g} s utterances in total, split contiguously over {nranks} GPUs "
                    f"({-(-args.total_utts // nranks)} per GPU at most), micro-batch {args.micro_batch}")
    else:
        workload = (f"cfg5 shard: {args.utts_per_gpu} x {args.seconds:g} s utterances per GPU "
                    f"({args.utts_per_gpu * nranks} over {nranks} GPUs), micro-batch {args.micro_batch}")
    scaling = "strong" if args.total_utts > 0 else "weak"

    if args.impl == "reference":
        if rank != 0:
            return
        v, sec, cores, kind = cpu_reference(args, args.cpu_utts, args.steps, args.warmup)
        sample = (f"{args.cpu_utts} x {args.seconds:g} s utterances per step, batch 1 each, fp32, torch CPU ({cores} threads), "
                  + ("the reference's own dataset.mel_spectrogram + Models.HiFiGAN (baseline/_ref)" if kind == "reference"
                     else "oracle port of the reference's calls (baseline/_ref not staged)"))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "audio-sec/sec", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "audio-sec/sec", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "audio-sec/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import synth
    import __graft_entry__ as entry
    pkg = entry.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its version banner to STDOUT (fd 1) when the communicator comes up: send fd 1 to stderr while it
        # does, so that stdout carries only the one JSON line the contract asks for
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = synth.HIFIGAN_V1
    h = synth.AttrDict(cfg)
    torch.manual_seed(cfg["seed"])
    gen = pkg.HiFiGAN(h)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state(cfg, 1234, "init").items()})
    gen = gen.to(dev).eval()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # the reference's "Removing weight norm..." print (hifigan.py:127)
        gen.remove_weight_norm()
    gen.precision = args.precision
    voc = pkg.Vocoder(gen, h, micro_batch=args.micro_batch, device=dev)

    U, T = args.utts_per_gpu, int(args.seconds * SR)
    if args.total_utts > 0:
        U = len(pkg.shard_range(args.total_utts, world, rank))
        if U == 0:
            raise SystemExit(f"--total-utts {args.total_utts} leaves rank {rank} of {world} without work")
    # every rank vocodes its own shard of the global utterance list (per-utterance sharding)
    wav_host = torch.from_numpy(synth.make_wave(U, T, 1000 + rank)).pin_memory()
    wav_dev = wav_host.to(dev)
    frames = 1 + T // cfg["hop_size"]
    t_out = frames * 256
    out_dev = torch.empty((U, t_out), dtype=torch.float32, device=dev)
    out_host = torch.empty((U, t_out), dtype=torch.float32).pin_memory()
    audio_sec_per_step = U * t_out / SR                     # this rank
    total_audio_sec_per_step = (args.total_utts if args.total_utts > 0 else U * world) * t_out / SR   # whole job

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # the front-end kernel by itself (BASELINE.json's second metric: mel front-end GB/s vs HBM peak), before the long step
    # heats the GPU into its power cap: the whole shard's waveforms in one launch, CUDA events, 20 launches
    with torch.no_grad():
        for _ in range(3):
            voc.mel(wav_dev)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(20):
            voc.mel(wav_dev)
        f1.record()
        torch.cuda.synchronize()
    fe_alone_ms = f0.elapsed_time(f1) / 20
    fe_alone_gbs = (4.0 * U * T + 4.0 * U * 80 * frames) / (fe_alone_ms * 1e-3) / 1e9

    dev_step = lambda: voc.run_device(wav_dev, out_dev)
    host_step = lambda: voc.run_host(wav_host, out_host)

    for _ in range(args.warmup):
        dev_step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = pkg._lib.launch_count()
    ms_total = timed(dev_step, args.steps)
    launches = pkg._lib.launch_count() - l0
    clocks = sampler.stop()
    if pkg._lib.tc_abort_status():
        raise SystemExit("tensor-core kernel tripped its bounded wait; results invalid")

    # The timed output is CHECKED: the first micro-batch of what the timed steps left in out_dev against the fp32
    # CUDA-core path of the same module on the same input (itself within 1e-4 of the reference, tests/test_gpu_parity.py);
    # de-meaned SNR as Metrics/snr.py:25-31, the gate of the 16-bit path is >= 40 dB.
    def snr_db(ref, deg):
        ref = ref.double() - ref.double().mean(dim=-1, keepdim=True)
        deg = deg.double() - deg.double().mean(dim=-1, keepdim=True)
        return float((10 * torch.log10(ref.pow(2).sum(-1) / (ref - deg).pow(2).sum(-1).clamp_min(1e-30))).min())

    n_chk = min(U, args.micro_batch)
    with torch.no_grad():
        gen.precision = "fp32"
        ref32 = gen(voc.mel(wav_dev[:n_chk]))
        gen.precision = args.precision
    parity = {"parity_snr_db": snr_db(ref32, out_dev[:n_chk]), "checked": f"{n_chk} utterances of the timed output vs the fp32 path",
              "gate_db": 40.0, "max_abs_err": float((ref32 - out_dev[:n_chk]).abs().max())}
    del ref32
    if not parity["parity_snr_db"] >= (40.0 if args.precision == "bf16" else 80.0):
        raise SystemExit(f"timed output fails the parity gate: {parity}")

    for _ in range(args.warmup):
        host_step()
    ms_e2e = timed(host_step, args.steps)
    if pkg._lib.tc_abort_status():
        raise SystemExit("tensor-core kernel tripped its bounded wait during the host leg; results invalid")
    parity["e2e_vs_device_max_abs"] = float((out_host[:n_chk].to(dev) - out_dev[:n_chk]).abs().max())
    if parity["e2e_vs_device_max_abs"] != 0.0:
        raise SystemExit(f"host-buffer path and device-resident path disagree: {parity}")

    # per-kernel event timing over the same K steps (separate pass so the events do not perturb `value`)
    pkg._lib.profile_begin()
    for _ in range(args.steps):
        dev_step()
    torch.cuda.synchronize()
    prof = pkg._lib.profile_end()

    value = total_audio_sec_per_step * args.steps / (ms_total / 1e3)
    e2e = total_audio_sec_per_step * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    # DRAM traffic of the tensor-core kernels from the committed `ncu --set full` capture of one forward
    # (profiles/r02_ncu_traffic.json: bytes per launch, same shapes as this workload's micro-batch)
    traffic, traffic_note = None, "no ncu capture for this micro-batch"
    tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        if tj.get("micro_batch") == args.micro_batch and tj.get("seconds") == args.seconds:
            traffic = tj["dram_bytes_per_launch"]
            traffic_note = (f"dram__bytes_read+write per launch, mean over the {tj['launches']} tensor-core launches of one forward "
                            f"(algorithmic bytes per launch by the same count: {tj['algorithmic_bytes_per_launch']:.3e})")
    tc = [k for k in prof if k["kernel"].startswith(("conv_tc", "ups_tc", "resblock_tc", "pair_tc"))]
    tc_ms, tc_flops, tc_n = sum(k["ms"] for k in tc), sum(k["flops"] for k in tc), sum(k["launches"] for k in tc)
    all_ms = sum(k["ms"] for k in prof)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else 0.0
    fe = [k for k in prof if k["kernel"].startswith("mel_frontend")]
    fe_gbs = sum(k["bytes"] for k in fe) / (sum(k["ms"] for k in fe) * 1e-3) / 1e9 if fe else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "audio-sec/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "parity": parity,
        "config": {"workload": workload, "weights": "random-init (reference-like N(0,0.01), seed 1234), weight_norm folded",
                   "l2": "inputs larger than L2 (113 MB waveforms, GB-scale activations per step)",
                   "operands": "16-bit tensor-core path: bf16 MRF convs, IEEE-half upsamplers and C<=32 second convs, fp32 accumulate, "
                               "fp32 residual stream / conv_pre / conv_post" if args.precision == "bf16" else "fp32 CUDA-core path",
                   "frames_per_utt": frames, "utts_this_rank": U, "flop_per_step_per_gpu": FLOP_PER_FRAME * frames * U},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "audio-sec/sec", "h2d_bytes_per_step": int(U * T * 4), "d2h_bytes_per_step": int(U * t_out * 4),
                "bytes_note": "per rank",
                "ms_per_step": ms_e2e / args.steps, "api": "Vocoder.run_host: pinned host wav -> mel_spectrogram -> HiFiGAN -> pinned host wav"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "resblock_tc / resblock_pipe / pair_tc + ups_tc / conv_tc kernels (tcgen05 implicit-GEMM convs: fused MRF ResBlocks, upsamplers, conv_pre)",
                     "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic, "traffic_note": traffic_note, "peak_source": pk["source"] + ", sustained bf16",
                     "launches": tc_n, "avg_launch_ms": tc_ms / tc_n if tc_n else None, "share_of_step": tc_ms / all_ms if all_ms else None,
                     "timing": "per-launch CUDA events on the launching stream, separate pass of the same K steps"},
        "frontend": {"kernel": "mel_frontend2_kernel (csrc/frontend.cu)", "achieved": fe_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": fe_gbs / pk["hbm_gbs"],
                     "achieved_alone": fe_alone_gbs, "frac_alone": fe_alone_gbs / pk["hbm_gbs"], "ms_alone": fe_alone_ms,
                     "note": "achieved = inside the timed step (per-launch events); achieved_alone = the kernel over the whole shard before the step",
                     "bound": "hbm (algorithmic bytes: waveform in + log-mel out)"},
        "kernels": sorted(({"kernel": k["kernel"], "launches": k["launches"], "ms_per_step": k["ms"] / args.steps,
                            "tflops": k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] else 0.0,
                            "gbs": k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] else 0.0} for k in prof),
                          key=lambda r: -r["ms_per_step"]),
    }
    extras = world == 1 and not args.no_extras
    if extras:  # secondary numbers: none of them may cost the headline line
        for key, fn in (("cfg1_gpu", lambda: cfg1_gpu(voc, synth, dev)), ("cfg2_frontend", lambda: cfg2_frontend(voc, synth, cfg, dev, pk)),
                        ("cfg3_hifigan", lambda: cfg3_hifigan(gen, synth, dev, pk, args.precision)),
                        ("cfg4_istftnet", lambda: cfg4_istftnet(pkg, synth, dev)),
                        ("eager_b200", lambda: eager_competitor(args, synth, cfg, dev, wav_dev)),
                        ("train_step", lambda: train_step(pkg, synth, cfg, dev)), ("disc_step", lambda: disc_step(pkg, dev))):
            try:
                line[key] = fn()
            except Exception as e:  # noqa: BLE001
                line[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores, kind = cpu_reference(args, args.cpu_utts, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "audio-sec/sec", "cores": cores, "kind": kind,
                                "sample": f"{args.cpu_utts} x {args.seconds:g} s utterances x 3 steps, batch 1 each, fp32 torch CPU"}
        if extras:
            v1, sec1, _, _ = cpu_reference(args, 1, 5, 1, seconds=2.0)
            line["cfg1_cpu"] = {"value": v1, "unit": "audio-sec/sec", "ms": sec1 * 1e3, "cores": cores, "kind": kind,
                                "config": "cfg1: HiFi-GAN V1, batch 1, 2 s waveform -> mel -> waveform, fp32 torch CPU"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
