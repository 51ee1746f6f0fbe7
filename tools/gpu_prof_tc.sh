#!/bin/bash
# ncu --set full capture of the tensor-core conv kernels (one GPU).  usage: tools/gpu_prof_tc.sh <tag>
set -u
OUT=gpurun_out/${1:-r01}
mkdir -p $OUT
for K in "32 1 1" "64 1 1" "128 1 1" "128 0 1"; do
  set -- $K
  N=tcp$1_$2_$3
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:tc_conv_persist_kernel<.int.$1, .bool.$2, .int.$3>" -s 40 -c 1 -o $OUT/prof_$N -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/prof_$N.log 2>&1
  tail -1 $OUT/prof_$N.log | cut -c1-100
done
ls -la $OUT | grep ncu-rep
