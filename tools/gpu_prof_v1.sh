#!/bin/bash
# ncu --set full of the AD v1 (HiFi-GAN vocoder) launches the round-1 verdict asked about: blocks.3.convs2.* / convs1.0 (grouped k11, 3 x 32
# channels: NT = 32 plain kernels with the LeakyReLU pre-activation), upsamples.* (plain NT = 128 / 64 with LeakyReLU).  usage: tools/gpu_prof_v1.sh <tag>
set -u
OUT=gpurun_out/${1:-prof_v1}
mkdir -p $OUT
BA="--workload v1 --no-extra --no-parity --regions 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_v1.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline $BA > $OUT/launches_v1.log 2>&1
python tools/launch_summary.py $OUT/launches_v1.csv 57 > $OUT/launch_summary_v1.txt 2>&1
tail -16 $OUT/launch_summary_v1.txt
cap() {
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o $OUT/prof_$1 -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline $BA > $OUT/prof_$1.log 2>&1
  if [ -f $OUT/prof_$1.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/prof_$1.ncu-rep > $OUT/ncu_full_v1_$1.txt 2>&1
    ncu -i $OUT/prof_$1.ncu-rep --page source --csv --print-source sass,cuda 2>/dev/null | python tools/ncu_lines.py 25 >> $OUT/ncu_full_v1_$1.txt 2>&1
    grep -E "gpu__time_duration.sum|pipe_tensor|issue_active|dram__bytes|lts__throughput" $OUT/ncu_full_v1_$1.txt | head -7
    rm -f $OUT/prof_$1.ncu-rep
  else
    echo "capture $1 failed"; tail -3 $OUT/prof_$1.log
  fi
}
# the decoder's NT=32 plain launches with pre-activation 2 (LeakyReLU): blocks.3.convs1.* / convs2.* come last in every step
cap nt32_lrelu_blocks3 "tc_conv_f16_kernel<.int.32, .bool.0, .int.2, .int.3>" 40
cap nt128_lrelu_upsample "tc_conv_f16_kernel<.int.128, .bool.0, .int.2, .int.3>" 20
cap nt64_lrelu "tc_conv_f16_kernel<.int.64, .bool.0, .int.2, .int.3>" 20
