#!/usr/bin/env python
"""What a user of the reference sees: the UNMODIFIED infers/inference_hifigan.py (per-file loop, batch 1, device forced to CPU by the
script) on a synthetic list of utterances, once against the reference's own modules on the host cores and once with the drop-in
directory on PYTHONPATH; reports the audio-seconds per second the script itself prints (infers/inference_hifigan.py:99-102).
Needs baseline/_ref (staged by __graft_entry__.build()) and tests/shims.  usage: script_bench.py [n_files=40]"""
import json, os, pathlib, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT]
import test_dropin_scripts as T  # noqa: E402  (harness helpers: synthetic corpus, checkpoint, environment)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_long = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # a longer list for the drop-in arm only: separates its one-time start-up from the per-file cost
rng = np.random.default_rng(1)
lengths = [int(v) for v in rng.integers(2 * 22050, 6 * 22050, size=max(n, n_long))]
with tempfile.TemporaryDirectory() as d:
    tmp = pathlib.Path(d)
    cfg, cfg_path, names = T._workdir(tmp, "hifigan_v1_config.json", lengths, seed=900, regime="unit")
    res = {}
    arms = [("reference (CPU, the script's own device)", False, {"CUDA_VISIBLE_DEVICES": ""}, n), ("drop-in on PYTHONPATH", True, {}, n)]
    if n_long > n:
        arms.append((f"drop-in on PYTHONPATH, {n_long} files", True, {}, n_long))
    flist_all = open(cfg["test_input_wavs_dir"]).read().splitlines(True)
    for arm, dropin, extra, count in arms:
        fl = tmp / f"list_{count}.txt"
        fl.write_text("".join(flist_all[:count]))
        c = dict(cfg, test_output_dir=str(tmp / f"out_{dropin}_{count}"), test_input_wavs_dir=str(fl))
        cp = tmp / f"cfg_{dropin}_{count}.json"
        cp.write_text(json.dumps(c))
        p = T._run([os.path.join(T.REF, "infers", "inference_hifigan.py"), "--cfg_filename", str(cp)], cwd=os.path.join(T.REF, "infers"),
                   env=T._env(dropin, extra), timeout=1800)
        nums = [float(x) for x in p.stdout.split() if x.replace(".", "", 1).replace("e-", "", 1).isdigit()]
        res[arm] = nums[-3:]
        print(f"{arm}: {nums[-2]:.1f} audio-s in {nums[-3]:.3f} s = {nums[-1]:.1f} audio-s/s ({count} files of 2 .. 6 s, batch 1 each, incl. file I/O)")
    vals = list(res.values())
    print(f"speed-up seen by the script on the same {n} files: {vals[1][-1] / vals[0][-1]:.1f}x")
    if len(vals) > 2:
        per_file = (vals[2][0] - vals[1][0]) / (n_long - n)
        print(f"drop-in: {per_file * 1e3:.2f} ms per further file (start-up inside the script's timer: {vals[1][0] - n * per_file:.2f} s)")
