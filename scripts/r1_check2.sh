#!/bin/bash
# validation run: parity suite, default bench, inverse-NTT A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 0 1; do
  timeout 300 python bench.py --no-cpu-baseline --intt-variant $v > gpurun_out/bench_intt$v.json 2> gpurun_out/bench_intt$v.err
  cat gpurun_out/bench_intt$v.json
done
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
