#!/bin/bash
# one dev iteration on the GPU: layer + parity tests (fail fast), headline bench with breakdown, optional stream_v1 breakdown.  usage: tools/gpu_step.sh <tag> [stream]
set -u
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee $OUT/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-extra > $OUT/bench.json 2> $OUT/bench.err
grep -E "res_units.0 |blocks.[0-3].conv |sum of|res_units.0.conv2 |conv2 |encoder.conv |conv1 |project" $OUT/bench.err | awk '{printf "%s %s | ", $1, $2} END {print ""}'
python - <<PY
import json
d=json.load(open('$OUT/bench.json')); p=d.get('parity',{})
print('  step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity idx_equal', p.get('idx_equal'), 'frames_differing', p.get('frames_differing'), 'wave', p.get('wave_max_abs'), d['clocks'].get('reasons'), d['clocks'].get('power_w_max'))
PY
if [ "${2:-}" = stream ]; then
  timeout 600 python bench.py --workload stream_v1 --steps 20 --warmup 3 --breakdown --no-cpu-baseline --no-extra --no-parity > $OUT/stream.json 2> $OUT/stream.err
  grep -E " ms " $OUT/stream.err | awk '{printf "%s %s | ", $1, $2} END {print ""}'
  python -c "
import json; d=json.load(open('$OUT/stream.json')); print('  stream step', round(d['ms_per_step'],3), 'launches', d.get('gpu_launches'))"
fi
