for E in "ADEC_GSPAN=0" "ADEC_GSPAN=1"; do
  env $E timeout 200 python bench.py --workload v1_bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-parity --regions 3 --breakdown > gpurun_out/gsbf_$E.json 2> gpurun_out/gsbf_$E.err
  python -c "
import json; d=json.load(open('gpurun_out/gsbf_$E.json')); print('$E v1_bf16', round(d['ms_per_step'],3), d['timed_regions_ms_per_step'])"
  grep -E "blocks.[0-3].convs1.0 |blocks.[0-3].convs2.0 |upsamples|conv_out|sum of" gpurun_out/gsbf_$E.err | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
