#!/bin/bash
# headline bench (with the oracle parity block) under different environment settings.  usage: tools/gpu_env.sh <tag> "<ENV=V ...>" ...
set -u
OUT=gpurun_out/$1
shift
mkdir -p $OUT
i=0
for E in "$@"; do
  i=$((i+1))
  echo "== $E"
  env $E timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-extra > $OUT/env_$i.json 2> $OUT/env_$i.err
  grep -E "res_units.0 |blocks.[0-3].conv |sum of|res_units.0.conv2 |conv2 |encoder.conv |conv1 |project" $OUT/env_$i.err | awk '{printf "%s %s | ", $1, $2} END {print ""}'
  python - <<PY
import json
try:
    d=json.load(open('$OUT/env_$i.json')); p=d.get('parity',{})
    print('  step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity idx_equal', p.get('idx_equal'), 'frames_differing', p.get('frames_differing'), 'wave', p.get('wave_max_abs'), 'ties', p.get('tie_margins'), d['clocks'].get('reasons'), d['clocks'].get('power_w_max'))
except Exception as e:
    print('  failed:', e); print(open('$OUT/env_$i.err').read()[-600:])
PY
done
