#!/bin/bash
mkdir -p gpurun_out
for w in fe 48k finetune; do
  timeout 200 python bench.py --workload $w --model tiny --steps 2 --warmup 1 > gpurun_out/r2m_tiny_$w.json 2> gpurun_out/r2m_tiny_$w.err || { echo "tiny $w FAILED"; tail -5 gpurun_out/r2m_tiny_$w.err; }
  python -c "
import json;d=json.load(open('gpurun_out/r2m_tiny_$w.json'));print('tiny $w', round(d['value'],1), 'clips/s')" 2>/dev/null
done
timeout 300 python bench.py --workload fe > gpurun_out/r2m_fe.json 2> gpurun_out/r2m_fe.err; tail -c 900 gpurun_out/r2m_fe.json; tail -2 gpurun_out/r2m_fe.err
timeout 400 python bench.py --workload 48k --steps 4 > gpurun_out/r2m_48k.json 2> gpurun_out/r2m_48k.err; tail -c 900 gpurun_out/r2m_48k.json; tail -2 gpurun_out/r2m_48k.err
timeout 400 python bench.py --workload finetune > gpurun_out/r2m_finetune.json 2> gpurun_out/r2m_finetune.err; tail -c 900 gpurun_out/r2m_finetune.json; tail -2 gpurun_out/r2m_finetune.err
timeout 300 python bench.py --workload finetune --frozen > gpurun_out/r2m_finetune_frozen.json 2> gpurun_out/r2m_finetune_frozen.err; tail -c 600 gpurun_out/r2m_finetune_frozen.json; tail -2 gpurun_out/r2m_finetune_frozen.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2m_bench.json'));print('pretrain', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['e2e']['value'], d['roofline']['frac'])"; tail -2 gpurun_out/r2m_bench.err
