#!/bin/bash
# quick GPU pass: all gpu tests + headline bench without baselines.  usage: tools/gpu_quick.sh <tag> [extra bench args]
set -u
OUT=gpurun_out/${1:-r02}
shift || true
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline "$@" > $OUT/bench_quick.json 2> $OUT/bench_quick.err
grep -E "res_units.0 |blocks.[0-3].conv |sum of|conv2 " $OUT/bench_quick.err | head -30
python - <<PY
import json
d=json.load(open('$OUT/bench_quick.json'))
print('value', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step'],3), d['clocks'])
print('parity', d.get('parity'))
for k,v in d.get('extra_workloads',{}).items(): print(k, json.dumps(v)[:1200])
PY
