#!/usr/bin/env python
"""cProfile of the UNMODIFIED infers/inference_hifigan.py with the drop-in on PYTHONPATH (200 files of 2 .. 6 s): where the per-file
milliseconds of the script path go.  usage: script_profile.py [n_files=200]"""
import json, os, pathlib, pstats, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT]
import test_dropin_scripts as T  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(1)
lengths = [int(v) for v in rng.integers(2 * 22050, 6 * 22050, size=n)]
with tempfile.TemporaryDirectory() as d:
    tmp = pathlib.Path(d)
    cfg, cfg_path, names = T._workdir(tmp, "hifigan_v1_config.json", lengths, seed=900, regime="unit")
    prof = str(tmp / "script.prof")
    p = subprocess.run([sys.executable, "-m", "cProfile", "-o", prof, os.path.join(T.REF, "infers", "inference_hifigan.py"), "--cfg_filename", cfg_path],
                       cwd=os.path.join(T.REF, "infers"), env=T._env(True), capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    print(p.stdout.split()[-3:])
    st = pstats.Stats(prof)
    st.sort_stats("cumulative").print_stats(45)
