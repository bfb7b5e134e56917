#!/bin/bash
# usage: tools/bench_kernels.sh <tag> [extra bench args]; prints the per-kernel table of one short bench run
tag=$1; shift
python bench.py --steps 2 --warmup 2 --utts-per-gpu 32 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python - "$tag" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(f"== {sys.argv[1]}: value {d['value']:.0f} audio-s/s  ms/step {d['ms_per_step']:.1f}  conv_tc {d['roofline']['achieved']:.0f} TF/s ({100*d['roofline']['frac']:.1f}%)  e2e {d['e2e']['value']:.0f}")
for k in d["kernels"]: print(f"   {k['kernel']:22s} x{k['launches']:4d} {k['ms_per_step']:8.2f} ms  {k['tflops']:7.1f} TF/s {k['gbs']:7.0f} GB/s")
PY
