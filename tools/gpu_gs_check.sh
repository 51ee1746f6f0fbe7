#!/bin/bash
# span-mode check at batch scale: golden tests, then v1_bf16 (bf16 NT=128 launches use span mode) and the headline with ADEC_GSPAN=1, each under a timeout
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "whole_piece or bf16 or oneshot" 2>&1 | tail -2
bash tools/gpu_gs_bf16.sh 2>&1 | grep -v "^Traceback\|^  File\|^    \|json" | head -8
timeout 300 bash tools/gpu_env.sh gs4 "ADEC_GSPAN=1" 2>&1 | grep -E "^==|step|failed"
