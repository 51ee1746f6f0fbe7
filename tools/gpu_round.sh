#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, ncu launch list, ncu full captures of the top kernels.
# usage: tools/gpu_round.sh <tag> [full]
set -u
TAG=${1:-r01}
MODE=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches.log 2>&1
if [ "$MODE" = "full" ]; then
  for K in "32 32 256 32 1" "64 64 128 32 1" "128 128 64 16 1" "256 256 64 8 1"; do
    set -- $K
    N=ru$1
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
        -k "regex:conv_gemm_kernel<.int.$1, .int.$2, .int.$3, .int.$4, .bool.$5>" -s 40 -c 2 -o $OUT/prof_$N -f \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/prof_$N.log 2>&1
  done
fi
ls -la $OUT
