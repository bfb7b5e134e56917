#!/bin/bash
# fused ResBlock kernel over the HiFi-GAN V1 MRF shapes at cfg5 micro-batch size (32 x 862 frames)
for k in 3 7 11; do python tools/rb_bench.py 32 $k 220672 32 3; done
for k in 3 7 11; do python tools/rb_bench.py 64 $k 110336 32 3; done
for k in 3 7 11; do python tools/rb_bench.py 128 $k 55168 32 1; done
for k in 3 7 11; do python tools/rb_bench.py 256 $k 6896 32 1; done
