#!/usr/bin/env python
"""cProfile (by internal time) of the UNMODIFIED train_time_wi_inv.py on the drop-in, 32 steps: where a step's host time goes.
usage: train_script_profile.py [extra env as K=V ...]"""
import json, os, pathlib, pstats, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT]
import test_dropin_scripts as T  # noqa: E402
extra = dict(kv.split("=", 1) for kv in sys.argv[1:])
n_utts, batch = 256, 16
with tempfile.TemporaryDirectory() as d:
    tmp = pathlib.Path(d)
    cfg, cfg_path, names = T._workdir(tmp, "hifigan_v1_config.json", [22050] * n_utts + [6000], seed=700, batch_size=batch, segment_size=8192,
                                      num_workers=2, checkpoint_interval=100000, summary_interval=100000, stdout_interval=100000,
                                      validation_interval=100000, training_epochs=2)
    (tmp / "val.txt").write_text(f"DUMMY1/{names[-1]}|v\n")
    (tmp / "train.txt").write_text("".join(f"DUMMY1/{n}|t\n" for n in names[:n_utts]))
    c = dict(cfg, checkpoint_path=str(tmp / "ck"), input_training_wav_list=str(tmp / "train.txt"), input_validation_wav_list=str(tmp / "val.txt"))
    cp = tmp / "train_cfg.json"
    cp.write_text(json.dumps(c))
    prof = str(tmp / "train.prof")
    p = subprocess.run([sys.executable, "-m", "cProfile", "-o", prof, os.path.join(T.REF, "train_time_wi_inv.py"), "--cfg_filename", str(cp)],
                       cwd=T.REF, env=T._env(True, extra), capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    pstats.Stats(prof).sort_stats("tottime").print_stats(28)
