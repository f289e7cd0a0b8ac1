#!/bin/bash
# two GPUs: the world-2 trainer tests over NCCL (bf16 and fp32 gradient buckets) and the 2-rank bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_ddp_cpu.py -q -m "gpu or not gpu" 2>&1 | tail -15 > gpurun_out/n2_tests.log
grep -E "passed|failed|FAILED|^E  |skipped" gpurun_out/n2_tests.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -c 1500 gpurun_out/n2_bench.json; tail -5 gpurun_out/n2_bench.err
python -c "
import json
for line in open('gpurun_out/n2_bench.json'):
    if line.startswith('{'):
        d=json.loads(line); print('N2', d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'])"
