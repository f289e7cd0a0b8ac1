#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "slab" 2>&1 | tail -5 > gpurun_out/r2ag_slab.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2ag_slab.log | head
TAG=r2ag bash tools/gpu_all.sh
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2ag_profile_step.txt 2>&1; grep -E "step |slab|relayout" gpurun_out/r2ag_profile_step.txt
