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
echo "== quick parity of the relaxed-range transforms and the new fold kernel"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "ntt or fold or process_query or expansion or pack" 2>&1 | tail -3
echo "== first dimension: analysis knobs"
run "default (ksps 4)" X=1 --
run "no MMA" B200PIR_TC5_DBG=1 --
run "no epilogue" B200PIR_TC5_DBG=2 --
run "copy only" B200PIR_TC5_DBG=3 --
run "ksps 2 (8 KiB x 12)" B200PIR_TC5_KSPS=2 --
run "ksps 8 (32 KiB x 3)" B200PIR_TC5_KSPS=8 --
run "ksps 1 (4 KiB x 24)" B200PIR_TC5_KSPS=1 --
echo "== fold variants"
run "fold old 3/SM" X=1 -- --fold-variant 1
run "fold old 2/SM" X=1 -- --fold-variant 0
run "fold lz 3/SM" X=1 -- --fold-variant 2
run "fold lz 2/SM" X=1 -- --fold-variant 3
} 2>&1 | tee gpurun_out/gpu_exp.log
