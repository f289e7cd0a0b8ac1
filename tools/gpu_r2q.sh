#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_parity_gaps_gpu.py tests/test_dropin_gpu.py tests/test_finetune_gpu.py -q -m gpu -k "gathered or strided or step_ or large_config or stochastic or dropin or finetune" 2>&1 | tail -40 > gpurun_out/r2q_tests.log; echo "exit $?" >> gpurun_out/r2q_tests.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2q_tests.log | head -30
timeout 300 python bench.py --no-cpu-baseline --batch 32 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2q_bench.json'));print('pretrain B=32', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['e2e']['value'], d['mem_gb'])"; tail -2 gpurun_out/r2q_bench.err
B=24 timeout 300 python tools/profile_step.py > gpurun_out/r2q_profile_step.txt 2>&1; head -3 gpurun_out/r2q_profile_step.txt; grep -E "taps=19|rows=42624\]|row_gather|neigh" gpurun_out/r2q_profile_step.txt | head
