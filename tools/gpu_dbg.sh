#!/bin/bash
# timing experiments with the ADEC_DBG_* knobs (wrong results, timing only).  usage: tools/gpu_dbg.sh <tag> "<ENV=V ...>" ...
set -u
OUT=gpurun_out/$1
shift
mkdir -p $OUT
i=0
for E in "$@"; do
  i=$((i+1))
  echo "== $E"
  env $E timeout 300 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline --no-extra --no-parity > $OUT/dbg_$i.json 2> $OUT/dbg_$i.err
  grep -E "res_units.0 |blocks.[0-3].conv |sum of|res_units.0.conv2 |conv2 |encoder.conv |conv1 |project" $OUT/dbg_$i.err | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
