#!/bin/bash
# One GPU-box session: the whole `-m gpu` suite, the default bench line, and an ncu capture of the first-dimension kernel.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-check}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_${TAG}.log 2>&1
grep -E "^(FAILED|ERROR)|AssertionError|passed|failed" gpurun_out/pytest_${TAG}.log | cut -c1-400 | tail -15
echo "== bench (default)"
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % tag).read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "kernel_ms",
          round(d["roofline"]["kernel_ms"], 3), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
    print("single", d["single_query_latency_ms"], "sweep", d.get("concurrent_queries_sweep"))
    print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "extrapolated_sample"})
except Exception as e:
    print("bench failed:", e, open("gpurun_out/bench_%s.err" % tag).read()[-1500:])
PY
echo "== ncu: first-dimension kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_multiply_tc5 -s 3 -c 1 -o gpurun_out/ncu_tc5_${TAG} -f \
  python bench.py --steps-only --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_tc5_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_tc5_${TAG}.log | cut -c1-300
} 2>&1 | tee gpurun_out/gpu_check_${TAG}.log
