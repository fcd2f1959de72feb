#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "imma or batch or three_phase or sharded or fold or pack or expan" > gpurun_out/pytest_gpu_16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_16.log
tail -5 gpurun_out/pytest_gpu_16.log
for v in 16 8; do
  timeout 300 python bench.py --no-cpu-baseline --queries-per-pass $v > gpurun_out/bench_qpp$v.json 2> gpurun_out/bench_qpp$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_qpp$v.json"))
print("queries_per_pass $v", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["stage_ms_per_step"])
PY
done
