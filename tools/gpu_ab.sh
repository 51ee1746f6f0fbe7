#!/bin/bash
# A/B bench of alternative builds.  usage: tools/gpu_ab.sh <tag> <lib1> [<lib2> ...]   (lib = path of a .so, "default" = in-tree build)
set -u
OUT=gpurun_out/$1
shift
mkdir -p $OUT
for LIB in "$@"; do
  N=$(basename $LIB .so)
  if [ "$LIB" = default ]; then unset ADEC_LIB_PATH; else export ADEC_LIB_PATH=$PWD/$LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-extra > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  echo "== $N"
  grep -E "res_units.0 |blocks.[0-3].conv |sum of|res_units.0.conv2 " $OUT/bench_$N.err | head -30
  python -c "
import json; d=json.load(open('$OUT/bench_$N.json')); print('$N', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['parity']['idx_equal'], d['parity']['wave_max_abs'], d['clocks']['power_w_max'])"
done
