#!/bin/bash
# One-box measurement run (round 1, second half): tests, bench lines, ncu evidence.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_final2.log; cat gpurun_out/pytest_gpu_final2.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench2_default.json 2> gpurun_out/bench2.err
timeout 300 python bench.py --queries-per-pass 16 --no-cpu-baseline > gpurun_out/bench2_qpp16.json 2>> gpurun_out/bench2.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench2_batch1.json 2>> gpurun_out/bench2.err
timeout 300 python bench.py --batch 4 --no-cpu-baseline > gpurun_out/bench2_batch4.json 2>> gpurun_out/bench2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench2_reference.json 2>> gpurun_out/bench2.err
for f in default qpp16 batch1 batch4 reference; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench2_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],2), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('single_query_latency_ms'), d.get('cpu_baseline',{}).get('value'))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 2 --warmup 3 --steps-only --no-cpu-baseline > gpurun_out/ncu_final2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_multiply_imma8 -s 3 -c 1 -o gpurun_out/prof_imma8_final2 -f python bench.py --steps 2 --warmup 3 --steps-only --no-cpu-baseline >> gpurun_out/ncu_final2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_expand_round_res -s 8 -c 1 -o gpurun_out/prof_expand_res_final2 -f python bench.py --steps 2 --warmup 3 --steps-only --no-cpu-baseline >> gpurun_out/ncu_final2.log 2>&1
timeout 300 python scripts/bench_kernels.py ntt > gpurun_out/kernels_final2.jsonl 2>> gpurun_out/ncu_final2.log; cut -c1-300 gpurun_out/kernels_final2.jsonl
ls gpurun_out | head -50
