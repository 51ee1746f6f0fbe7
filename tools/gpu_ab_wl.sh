#!/bin/bash
# A/B of alternative builds on the other workloads.  usage: tools/gpu_ab_wl.sh <tag> "<workload> ..." <lib1> [<lib2> ...]   (lib = path of a .so, "default" = in-tree build)
set -u
OUT=gpurun_out/$1; WLS=$2; shift 2
mkdir -p $OUT
for LIB in "$@"; do
  N=$(basename $LIB .so)
  if [ "$LIB" = default ]; then unset ADEC_LIB_PATH; else export ADEC_LIB_PATH=$PWD/$LIB; fi
  for WL in $WLS; do
    ST=5; [ $WL = stream_v1 ] && ST=20
    timeout 600 python bench.py --workload $WL --steps $ST --warmup 3 --no-cpu-baseline --no-extra --no-parity --regions 3 > $OUT/bench_${N}_$WL.json 2> $OUT/bench_${N}_$WL.err
    python -c "
import json; d=json.load(open('$OUT/bench_${N}_$WL.json')); print('$N $WL', round(d['ms_per_step'],3), d['timed_regions_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],3))" 2>&1 | tail -1
  done
done
