#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "pair" 2>&1 | tail -12 > gpurun_out/r2ah_pair.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2ah_pair.log | head
TAG=r2ah bash tools/gpu_all.sh
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2ah_profile_step.txt 2>&1; grep -E "step |dgelu|epi=0001|N=4096,K=1024" gpurun_out/r2ah_profile_step.txt
