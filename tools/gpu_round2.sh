#!/bin/bash
# full GPU pass: all gpu tests, smoke, bench (headline + extras), reference arm.  usage: tools/gpu_round2.sh <tag>
set -u
OUT=gpurun_out/${1:-r02}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
ADEC_BENCH_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 --breakdown > $OUT/bench.json 2> $OUT/bench.err
tail -60 $OUT/bench.err
python - <<PY
import json
d=json.load(open('$OUT/bench.json'))
print('value', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step'],3), d['clocks'])
print('roofline frac', d['roofline']['frac'], 'kernel_frac', d['roofline']['kernel_frac'], 'compute', {k:v for k,v in d['roofline']['compute'].items() if k!='probe'})
print('parity', d.get('parity'))
for k,v in d.get('extra_workloads',{}).items(): print(k, json.dumps(v)[:900])
print('cpu', json.dumps(d.get('cpu_baseline'))[:600])
print('eager', d.get('gpu_eager_baseline'))
PY
