#!/bin/bash
# usage: tools/gpu_timeline.sh <tag> <op name> ...
# first build the instrumented library here (it travels with gpurun):
#   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DADEC_TIMELINE \
#        -o audiodec_b200/lib/libaudiodec_b200_tl.so audiodec_b200/csrc/adec.cu
OUT=gpurun_out/$1; shift; mkdir -p $OUT
export ADEC_LIB_PATH=$PWD/audiodec_b200/lib/libaudiodec_b200_tl.so
for OP in "$@"; do
  ADEC_TIMELINE_OP=$OP ADEC_TIMELINE_OUT=$OUT/tl_$OP.txt timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-parity > /dev/null 2> $OUT/tl_$OP.err
  echo "== $OP"; python tools/timeline_f16.py $OUT/tl_$OP.txt 40 24 | tail -34
done
