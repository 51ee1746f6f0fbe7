#!/bin/bash
# GPU check of the kind::f16 engine: layer tests, parity tests, bench with per-launch breakdown
set -u
OUT=gpurun_out/${1:-r02}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_layers_gpu.py -q 2>&1 | tail -8 | tee $OUT/pytest_layers.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -k "not tf32 and not ffma" 2>&1 | grep -E "parity\]|passed|failed|Error|error|assert" | tail -40 | tee $OUT/pytest_parity_f16.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown > $OUT/bench_f16.json 2> $OUT/bench_f16.err
python -c "import json,sys; d=json.load(open('$OUT/bench_f16.json')); print('f16:', round(d['ms_per_step'],3), 'ms/step', round(d['value']/1e6,1), 'M samples/s; e2e', round(d['e2e']['ms_per_step'],3), d['clocks'])"
cat $OUT/bench_f16.err | tail -50
