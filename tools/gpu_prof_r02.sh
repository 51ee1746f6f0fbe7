#!/bin/bash
# round-2 profile set on one GPU: ncu launch list of a step + `--set full` captures of the dominant kernels, summarised on the box
# (key metrics + stall reasons + per-source-line instruction / sample shares).  usage: tools/gpu_prof_r02.sh <tag>
set -u
OUT=gpurun_out/${1:-prof_r02}
mkdir -p $OUT
BA="--no-extra --no-parity --regions 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline $BA > $OUT/launches.log 2>&1
python tools/launch_summary.py $OUT/launches.csv 45 > $OUT/launch_summary_f16.txt 2>&1
tail -12 $OUT/launch_summary_f16.txt
cap() {   # name, kernel regex, skip
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o $OUT/prof_$1 -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline $BA > $OUT/prof_$1.log 2>&1
  if [ -f $OUT/prof_$1.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/prof_$1.ncu-rep > $OUT/ncu_full_$1.txt 2>&1
    ncu -i $OUT/prof_$1.ncu-rep --page source --csv --print-source sass,cuda 2>/dev/null | python tools/ncu_lines.py 30 >> $OUT/ncu_full_$1.txt 2>&1
    grep -E "gpu__time_duration.sum|pipe_tensor|issue_active|dram__bytes|lts__throughput|xbar2l1tex|l1tex__throughput" $OUT/ncu_full_$1.txt | head -9
    [ "${KEEP_REP:-}" = "$1" ] || rm -f $OUT/prof_$1.ncu-rep
  else
    echo "capture $1 failed"; tail -3 $OUT/prof_$1.log
  fi
}
cap tcf32_1_1_3   "tc_conv_f16_kernel<.int.32, .bool.1, .int.1, .int.3>" 12
cap tcf64_1_1_3   "tc_conv_f16_kernel<.int.64, .bool.1, .int.1, .int.3>" 12
cap tcf128_1_1_3  "tc_conv_f16_kernel<.int.128, .bool.1, .int.1, .int.3>" 12
cap tcf128_0_1_3_k7 "tc_conv_f16_kernel<.int.128, .bool.0, .int.1, .int.3>" 24
cap tcf128_0_0_3  "tc_conv_f16_kernel<.int.128, .bool.0, .int.0, .int.3>" 16
cap rvq           "rvq_kernel" 6
ls -la $OUT
