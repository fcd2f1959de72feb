#!/bin/bash
# One-box measurement run for the round: tests, bench lines, ncu evidence.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --batch 1 --no-cpu-baseline > gpurun_out/bench_batch1_imma.json 2>> gpurun_out/bench_default.err
python bench.py --batch 1 --db-format 0 --no-cpu-baseline > gpurun_out/bench_batch1_imad.json 2>> gpurun_out/bench_default.err
python bench.py --batch 4 --no-cpu-baseline > gpurun_out/bench_batch4_imma.json 2>> gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err
for f in default batch1_imma batch1_imad batch4_imma reference; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],2), round(d['e2e']['value'],2), d.get('roofline',{}).get('frac'), d.get('single_query_latency_ms'), d.get('cpu_baseline',{}).get('value'))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_multiply_imma -s 3 -c 1 -o gpurun_out/prof_imma_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_final.log 2>&1
python scripts/bench_kernels.py ntt dpir > gpurun_out/kernels_final.jsonl 2>> gpurun_out/ncu_final.log; cut -c1-400 gpurun_out/kernels_final.jsonl
ls gpurun_out
