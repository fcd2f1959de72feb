#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "imma or three_phase or sharded or batch" > gpurun_out/pytest_gpu_imma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_imma.log
tail -3 gpurun_out/pytest_gpu_imma.log
for v in 0 1; do
  timeout 300 python bench.py --no-cpu-baseline --imma-variant $v > gpurun_out/bench_imma$v.json 2> gpurun_out/bench_imma$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_imma$v.json"))
print("imma_variant $v", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["stage_ms_per_step"])
PY
done
