#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "rowln or resln or residual" 2>&1 | tail -25 > gpurun_out/r2an_first.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2an_first.log | head -20
if grep -q "failed\|FAILED\|rror" gpurun_out/r2an_first.log; then cat gpurun_out/r2an_first.log; exit 0; fi
TAG=r2an bash tools/gpu_all.sh
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2an_profile_step.txt 2>&1; grep -E "step |rowln_bwd" gpurun_out/r2an_profile_step.txt
