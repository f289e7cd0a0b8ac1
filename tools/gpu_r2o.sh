#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
( timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_datapipe_gpu.py tests/test_engine_gpu.py tests/test_trainer_gpu.py tests/test_finetune_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2o_tests.log; echo "exit $?" >> gpurun_out/r2o_tests.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2o_tests.log | head -30
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2o_bench.json'));print('pretrain', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['e2e']['value'], d['mem_gb'])"; tail -2 gpurun_out/r2o_bench.err
timeout 300 python bench.py --workload fe > gpurun_out/r2o_fe.json 2> gpurun_out/r2o_fe.err; python -c "
import json;d=json.load(open('gpurun_out/r2o_fe.json'));print('fe', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['mem_gb'])"; tail -2 gpurun_out/r2o_fe.err
B=24 timeout 300 python tools/profile_step.py > gpurun_out/r2o_profile_step.txt 2>&1; head -4 gpurun_out/r2o_profile_step.txt
