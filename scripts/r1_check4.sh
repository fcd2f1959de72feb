#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "expan or process_query or three_phase or wire or e2e" > gpurun_out/pytest_gpu_exp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_exp.log
tail -5 gpurun_out/pytest_gpu_exp.log
for v in 0 1; do
  timeout 300 python bench.py --no-cpu-baseline --expand-variant $v > gpurun_out/bench_exp$v.json 2> gpurun_out/bench_exp$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_exp$v.json"))
print("expand_variant $v", d["value"], d["e2e"]["value"], d["single_query_latency_ms"], d["stage_ms_per_step"])
PY
done
