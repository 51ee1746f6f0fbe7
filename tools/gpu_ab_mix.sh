#!/bin/bash
# usage: tools/gpu_ab_mix.sh <tag> <lib> ...  : headline + stream_v1 (with ADEC_PLAIN_TEAMS 2 and 1) per build
OUT=gpurun_out/$1; shift; mkdir -p $OUT
for LIB in "$@"; do
  N=$(basename $LIB .so)
  if [ "$LIB" = default ]; then unset ADEC_LIB_PATH; else export ADEC_LIB_PATH=$PWD/$LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-parity > $OUT/h_$N.json 2> $OUT/h_$N.err
  python -c "
import json; d=json.load(open('$OUT/h_$N.json')); print('$N headline', round(d['ms_per_step'],3), d['timed_regions_ms_per_step'])"
  for T in 2 1; do
    ADEC_PLAIN_TEAMS=$T timeout 600 python bench.py --workload stream_v1 --steps 20 --warmup 3 --no-cpu-baseline --no-extra --no-parity --regions 3 > $OUT/s_${N}_$T.json 2> $OUT/s_${N}_$T.err
    python -c "
import json; d=json.load(open('$OUT/s_${N}_$T.json')); print('$N stream teams=$T', round(d['ms_per_step'],3), d['timed_regions_ms_per_step'])"
  done
done
