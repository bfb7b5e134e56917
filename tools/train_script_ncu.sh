#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of a window of the UNMODIFIED trainer on the drop-in: which kernels a step is made of
# usage: tools/train_script_ncu.sh <out.csv> [env assignments...]
out=$1; shift
python - "$out" "$@" <<'PY'
import json, os, pathlib, subprocess, sys, tempfile
ROOT = os.getcwd()
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT]
import test_dropin_scripts as T
out = os.path.abspath(sys.argv[1])
extra = dict(kv.split("=", 1) for kv in sys.argv[2:])
extra.setdefault("NVSE_B200_NO_ANOMALY", "1")
with tempfile.TemporaryDirectory() as d:
    tmp = pathlib.Path(d)
    cfg, cfg_path, names = T._workdir(tmp, "hifigan_v1_config.json", [22050] * 128 + [6000], seed=700, batch_size=16, segment_size=8192,
                                      num_workers=0, checkpoint_interval=100000, summary_interval=100000, stdout_interval=100000,
                                      validation_interval=100000, training_epochs=1)
    (tmp / "val.txt").write_text(f"DUMMY1/{names[-1]}|v\n")
    (tmp / "train.txt").write_text("".join(f"DUMMY1/{n}|t\n" for n in names[:128]))
    c = dict(cfg, checkpoint_path=str(tmp / "ck"), input_training_wav_list=str(tmp / "train.txt"), input_validation_wav_list=str(tmp / "val.txt"))
    cp = tmp / "train_cfg.json"
    cp.write_text(json.dumps(c))
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--launch-skip", "6000", "-c", "2500", "--csv", "--log-file", out,
           sys.executable, os.path.join(T.REF, "train_time_wi_inv.py"), "--cfg_filename", str(cp)]
    p = subprocess.run(cmd, cwd=T.REF, env=T._env(True, extra), capture_output=True, text=True, timeout=1500)
    print(p.returncode, p.stderr[-500:])
PY
