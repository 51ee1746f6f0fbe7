#!/bin/bash
# ncu launch list of one bench step.  usage: tools/gpu_launches.sh <tag>
OUT=gpurun_out/${1:-r01}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches.log 2>&1
python tools/launch_summary.py $OUT/launches.csv ${2:-44}
