#!/bin/bash
# ncu --set full capture of the four fused residual-unit kernels (one GPU).  usage: tools/gpu_prof.sh <tag>
set -u
OUT=gpurun_out/${1:-r01}
mkdir -p $OUT
for K in "32 32 256 32 1" "64 64 128 32 1" "128 128 64 16 1" "256 256 64 8 1"; do
  set -- $K
  N=ru$1
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:conv_gemm_kernel<.int.$1, .int.$2, .int.$3, .int.$4, .bool.$5>" -s 40 -c 2 -o $OUT/prof_$N -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/prof_$N.log 2>&1
  tail -3 $OUT/prof_$N.log
done
ls -la $OUT
