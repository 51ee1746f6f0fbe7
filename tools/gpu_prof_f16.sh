#!/bin/bash
# ncu --set full capture of kind::f16 conv kernels (one GPU).  usage: tools/gpu_prof_f16.sh <tag> "<NT> <fuse> <pre> <prec> <skip>" ...
set -u
OUT=gpurun_out/$1
shift
mkdir -p $OUT
for K in "$@"; do
  set -- $K
  N=tcf$1_$2_$3_$4
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:tc_conv_f16_kernel<.int.$1, .bool.$2, .int.$3, .int.$4>" -s $5 -c 1 -o $OUT/prof_$N -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > $OUT/prof_$N.log 2>&1
  tail -1 $OUT/prof_$N.log | cut -c1-100
done
ls -la $OUT | grep ncu-rep
