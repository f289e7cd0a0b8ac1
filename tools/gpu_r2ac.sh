#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -15 > gpurun_out/r2ad_attn.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2ad_attn.log | head -20
TAG=r2ad bash tools/gpu_all.sh
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2ad_profile_step.txt 2>&1; head -12 gpurun_out/r2ad_profile_step.txt
