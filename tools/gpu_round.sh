#!/bin/bash
# Run on the GPU box (via gpurun): tests, smoke, bench (ours + reference arm), ncu launch list.  usage: tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -c 4000 $OUT/bench.json
ADEC_CONV_PATH=ffma timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_ffma.json 2>> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches.log 2>&1
python tools/launch_summary.py $OUT/launches.csv 44 > $OUT/launch_summary.txt; tail -14 $OUT/launch_summary.txt
ls -la $OUT
