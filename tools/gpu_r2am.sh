#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "slab" 2>&1 | tail -25 > gpurun_out/r2am_first.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2am_first.log | head -20
if grep -q "failed\|FAILED\|rror" gpurun_out/r2am_first.log; then cat gpurun_out/r2am_first.log; exit 0; fi
TAG=r2am bash tools/gpu_all.sh
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2am_profile_step.txt 2>&1; grep -E "step |slab" gpurun_out/r2am_profile_step.txt
