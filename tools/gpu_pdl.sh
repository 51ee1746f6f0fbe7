#!/bin/bash
# A/B of programmatic dependent launch (ADEC_PDL): parity tests, then the default bench (extras included) with it on and off
mkdir -p gpurun_out/pdl
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
for P in 1 0 1; do
  ADEC_PDL=$P timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pdl/b_$P.json 2> gpurun_out/pdl/b_$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/pdl/b_$P.json')); e=d['extra_workloads']; l=e['latency_b1_1500']
print('PDL=$P step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity']['idx_equal'], d['parity']['wave_max_abs'],
      '| v1', round(e['v1']['ms_per_step'],2), 'v1_bf16', round(e['v1_bf16']['ms_per_step'],2), 'stream', round(e['stream_v1']['ms_per_step'],3), 'server', round(e['stream_v1']['server']['step_ms_mean_std'][0],3), e['stream_v1']['parity']['idx_equal'],
      '| latency enc/dec/hifigan', round(l['encoder_ms_mean_std'][0],3), round(l['decoder_symAD_ms_mean_std'][0],3), round(l['decoder_hifigan_v1_ms_mean_std'][0],3))
PY
done
