#!/bin/bash
# Build a variant of the library next to the product one for same-box A/B timing:
#   tools/ab_build.sh <tag> "<extra nvcc -D flags>"   ->  csrc/libnvse_b200_<tag>.so   (use with NVSE_LIB=...)
set -e
tag=$1; flags=$2
cd "$(dirname "$0")/../neural-vocoders-as-speech-enhancers_b200/csrc"
mkdir -p build_$tag
for f in core frontend conv_f32 istft conv_tc resblock_tc pair_tc ups_tc layers_tc tc_abort generator weights grad wgrad_tc generator_train disc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -cudart static $flags -c $f.cu -o build_$tag/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o libnvse_b200_$tag.so build_$tag/*.o
echo built libnvse_b200_$tag.so
