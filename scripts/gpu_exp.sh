#!/bin/bash
# Bottleneck experiments (one GPU-box session): first-dimension analysis knobs and fold / expansion variants.
mkdir -p gpurun_out
run() {  # label, env..., -- bench args
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps-only --no-cpu-baseline --steps 10 --warmup 3 "$@" > gpurun_out/exp.json 2> gpurun_out/exp.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/exp.json").read().strip().splitlines()[-1])
    print("%-34s %7.1f q/s  mul_kernel_ms %.3f  " % (sys.argv[1], d["value"], d["roofline"]["kernel_ms"]), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "failed:", e, open("gpurun_out/exp.err").read()[-800:])
PY
}
{
echo "== new tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "error_behaviour or dpir or fold" 2>&1 | tail -6
echo "== fold variants"
run "fold lz 3/SM" X=1 -- --fold-variant 2
run "fold lz 3/SM + L1 prefetch" X=1 -- --fold-variant 5
run "fold lz 2/SM + L1 prefetch" X=1 -- --fold-variant 6
run "fold lz 2/SM regtw + prefetch" X=1 -- --fold-variant 7
} 2>&1 | tee gpurun_out/gpu_exp.log
