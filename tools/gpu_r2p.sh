#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_parity_gaps_gpu.py -q -m gpu -k "attention or step_ or large_config" 2>&1 | tail -40 > gpurun_out/r2p_tests.log; echo "exit $?" >> gpurun_out/r2p_tests.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2p_tests.log | head -30
for b in 24 32; do timeout 300 python bench.py --no-cpu-baseline --batch $b > gpurun_out/r2p_bench_$b.json 2> gpurun_out/r2p_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2p_bench_$b.json'));print('pretrain B=$b', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['e2e']['value'], d['mem_gb'])"; tail -2 gpurun_out/r2p_bench.err; done
