#!/bin/bash
# round-2 evidence: the front-end alone, a 32-utterance bench line, ncu --set full of the shipped front-end kernel and of the
# dominant generator kernels (summaries written on the box; .ncu-rep files kept only while gpurun_out stays small)
set -x
mkdir -p gpurun_out
python tools/fe_bench.py > gpurun_out/r02b_fe_plain.txt 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mel_frontend2 -c 1 -o gpurun_out/r02b_fe_full -f python tools/fe_bench.py 2048 4.0 > gpurun_out/r02b_fe_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02b_fe_full.ncu-rep > gpurun_out/r02b_ncu_full_frontend.txt 2>&1
python tools/ncu_hot_sass.py gpurun_out/r02b_fe_full.ncu-rep 0 40 > gpurun_out/r02b_fe_hot_sass.txt 2>&1
python bench.py --steps 2 --warmup 3 --utts-per-gpu 32 --no-cpu-baseline > gpurun_out/r02b_bench32.json 2> gpurun_out/r02b_bench32.err
timeout 600 ncu --set full --clock-control none -k regex:'resblock_pipe|ups_tc|conv_post1|resblock_tc_kernel|pair_tc' --launch-skip 60 -c 26 -o /tmp/r02b_gen_full -f python bench.py --steps 1 --warmup 1 --utts-per-gpu 32 --no-cpu-baseline > gpurun_out/r02b_gen_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02b_gen_full.ncu-rep > gpurun_out/r02b_ncu_full_generator.txt 2>&1
du -sh gpurun_out
