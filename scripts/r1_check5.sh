#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "expan or process_query_bytes or three_phase or wire" > gpurun_out/pytest_gpu_exp2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_exp2.log
tail -5 gpurun_out/pytest_gpu_exp2.log
for v in 0 2; do
  timeout 300 python bench.py --no-cpu-baseline --expand-variant $v > gpurun_out/bench_expv$v.json 2> gpurun_out/bench_expv$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_expv$v.json"))
print("expand_variant $v", d["value"], d["e2e"]["value"], d["single_query_latency_ms"], d["stage_ms_per_step"])
PY
done
