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
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_tcgen05.py -x -q 2>&1 | tail -15
echo "== first dimension with converged issue warps"
run "ksps8 B1 A4 (default)" X=1 --
run "ksps8 B2 A4" B200PIR_TC5_BBUFS=2 --
run "ksps4 B1 A4" B200PIR_TC5_KSPS=4 --
run "ksps4 B2 A2" B200PIR_TC5_KSPS=4 B200PIR_TC5_BBUFS=2 B200PIR_TC5_ABUFS=2 --
run "no epilogue" B200PIR_TC5_DBG=2 --
run "copy only" B200PIR_TC5_DBG=3 --
} 2>&1 | tee gpurun_out/gpu_exp.log
