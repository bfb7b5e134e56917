#!/bin/bash
# same-box A/B of the chain kernels: product library vs csrc/libnvse_b200_$1.so; C = 64 / 32 shapes of cfg5 (micro-batch 32), T32 layout
tag=$1
export NVSE_RB_T32=1
for k in 3 7 11; do
  for lib in "" neural-vocoders-as-speech-enhancers_b200/csrc/libnvse_b200_$tag.so; do
    echo "lib=${lib:-product}"; NVSE_LIB=$lib python tools/rb_bench.py 64 $k 110336 32 3
    NVSE_LIB=$lib NVSE_RB_H16=1 python tools/rb_bench.py 32 $k 220672 32 3
  done
done
