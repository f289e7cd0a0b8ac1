#!/bin/bash
# final evidence of the round: smoke(), every bench workload (the default one with its CPU baseline), reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.load(open('gpurun_out/$1'))
print('$1', round(d['value'],2), d['unit'], 'ms/step', round(d['ms_per_step'],2), 'frac', d.get('step_tensor_frac'), 'e2e', round(d['e2e']['value'],2), 'roofline', round(d['roofline']['frac'],3) if d.get('roofline') else None, 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; }
timeout 300 python bench.py --workload fe --no-cpu-baseline > gpurun_out/r2_bench_fe.json 2> gpurun_out/r2f_fe.err; show r2_bench_fe.json; tail -2 gpurun_out/r2f_fe.err
timeout 400 python bench.py --workload 48k --steps 8 --no-cpu-baseline > gpurun_out/r2_bench_48k.json 2> gpurun_out/r2f_48k.err; show r2_bench_48k.json; tail -2 gpurun_out/r2f_48k.err
timeout 400 python bench.py --workload finetune --no-cpu-baseline > gpurun_out/r2_bench_finetune.json 2> gpurun_out/r2f_ft.err; show r2_bench_finetune.json; tail -2 gpurun_out/r2f_ft.err
timeout 300 python bench.py --workload finetune --frozen --no-cpu-baseline > gpurun_out/r2_bench_finetune_frozen.json 2> gpurun_out/r2f_ftf.err; show r2_bench_finetune_frozen.json; tail -2 gpurun_out/r2f_ftf.err
timeout 600 python bench.py > gpurun_out/r2_bench_pretrain.json 2> gpurun_out/r2f_pre.err; show r2_bench_pretrain.json; tail -2 gpurun_out/r2f_pre.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2f_ref.err; tail -c 700 gpurun_out/r2_bench_reference_arm.json; tail -2 gpurun_out/r2f_ref.err
