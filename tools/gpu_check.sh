#!/bin/bash
# bounded first (new kernels), then the whole suite, the default bench line and the feature-extractor workload
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "sinc" 2>&1 | tail -25 > gpurun_out/chk_first.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/chk_first.log | head -20
if grep -q "failed\|FAILED\|rror" gpurun_out/chk_first.log || ! grep -q passed gpurun_out/chk_first.log; then cat gpurun_out/chk_first.log; exit 0; fi
TAG=chk bash tools/gpu_all.sh
timeout 300 python bench.py --workload fe --no-cpu-baseline > gpurun_out/chk_fe.json 2> gpurun_out/chk_fe.err; python -c "
import json;d=json.load(open('gpurun_out/chk_fe.json'));print('fe', d['value'], d['ms_per_step'])"
B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/chk_profile_step.txt 2>&1; grep -E "step |sinc" gpurun_out/chk_profile_step.txt; grep -E "fe_layer|a2v_gemm.*B=24" gpurun_out/chk_profile_step.txt | head -3
