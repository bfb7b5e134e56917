#!/usr/bin/env python
"""Seconds per optimiser step of the UNMODIFIED train_time_wi_inv.py (HiFi-GAN V1, batch 16 x 8192 samples, generator + MPD + MSD,
data loader with 2 workers) on a synthetic corpus: the reference's own modules on stock PyTorch (its defaults: TF32 convolutions)
against the drop-in on PYTHONPATH with its knobs.  Two runs per arm with different epoch counts; the difference removes start-up.
usage: train_script_bench.py [steps_short=16] [steps_long=48]   (multiples of 16)"""
import json, os, pathlib, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT]
import test_dropin_scripts as T  # noqa: E402

s_short = int(sys.argv[1]) if len(sys.argv) > 1 else 16
s_long = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n_utts, batch = 256, 16   # 16 steps per epoch: the per-epoch cost (the DataLoader forks its workers anew) is spread over them
per_epoch = n_utts // batch
arms = [("reference modules, stock PyTorch (TF32 convolutions, its default)", False, {}),
        ("drop-in, defaults (generator + mel on this repo's fp32 kernels, reference discriminators)", True, {}),
        ("drop-in, NVSE_B200_TRAIN_PRECISION=bf16 (generator forward / backward on the tensor cores)", True, {"NVSE_B200_TRAIN_PRECISION": "bf16"}),
        ("drop-in, bf16 generator + NVSE_B200_DISCRIMINATORS=1 (whole step on this repo's kernels)", True,
         {"NVSE_B200_TRAIN_PRECISION": "bf16", "NVSE_B200_DISCRIMINATORS": "1"}),
        ("drop-in, defaults + NVSE_B200_NO_ANOMALY=1 (the script's global anomaly detection neutralised)", True, {"NVSE_B200_NO_ANOMALY": "1"}),
        ("drop-in, bf16 generator + NVSE_B200_NO_ANOMALY=1", True, {"NVSE_B200_TRAIN_PRECISION": "bf16", "NVSE_B200_NO_ANOMALY": "1"})]
if os.environ.get("TSB_ARMS"):
    arms = [arms[int(i)] for i in os.environ["TSB_ARMS"].split(",")]
with tempfile.TemporaryDirectory() as d:
    tmp = pathlib.Path(d)
    cfg, cfg_path, names = T._workdir(tmp, "hifigan_v1_config.json", [22050] * n_utts + [6000], seed=700, batch_size=batch, segment_size=8192,
                                      num_workers=2, checkpoint_interval=100000, summary_interval=100000, stdout_interval=100000,
                                      validation_interval=100000)
    (tmp / "val.txt").write_text(f"DUMMY1/{names[-1]}|v\n")
    (tmp / "train.txt").write_text("".join(f"DUMMY1/{n}|t\n" for n in names[:n_utts]))
    for label, dropin, extra in arms:
        ts = []
        for steps in (s_short, s_long):
            c = dict(cfg, checkpoint_path=str(tmp / f"ck_{len(ts)}_{abs(hash(label))}"), input_training_wav_list=str(tmp / "train.txt"),
                     input_validation_wav_list=str(tmp / "val.txt"), training_epochs=steps // per_epoch)
            cp = tmp / "train_cfg.json"
            cp.write_text(json.dumps(c))
            t0 = time.perf_counter()
            T._run([os.path.join(T.REF, "train_time_wi_inv.py"), "--cfg_filename", str(cp)], cwd=T.REF, env=T._env(dropin, extra), timeout=1800)
            ts.append(time.perf_counter() - t0)
        print(f"{label}: {(ts[1] - ts[0]) / (s_long - s_short) * 1e3:.1f} ms per step  (runs of {s_short} / {s_long} steps: {ts[0]:.1f} / {ts[1]:.1f} s)", flush=True)
