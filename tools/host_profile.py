import os, sys, time, cProfile, pstats, io
import torch
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from util import pkg, build_generator
import synth
cfg = synth.HIFIGAN_V1
gen = build_generator(cfg, synth.make_state(cfg, 1234, "init"), "cuda", remove_wn=True)
gen.precision = "bf16"
a = cfg
y = torch.from_numpy(synth.make_wave(1, 44100, 3)).cuda()
def call():
    with torch.no_grad():
        m = pkg.mel_spectrogram(y, a["n_fft"], a["num_mels"], a["sampling_rate"], a["hop_size"], a["win_size"], a["fmin"], a["fmax"])
        return gen(m)
for _ in range(20): call()
torch.cuda.synchronize()
# host enqueue time (no sync) vs device time
t0 = time.perf_counter()
for _ in range(200): call()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/200:.3f} ms/call, total incl. drain {1e3*(t2-t0)/200:.3f} ms/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): call()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:5000])
