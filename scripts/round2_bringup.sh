#!/bin/bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget ended, in one box session,
# every step under its own timeout (the tcgen05 kernel's waits are bounded, a protocol error traps).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_bringup.sh'
mkdir -p gpurun_out
{
echo "== new ungated parity tests (golden replay, file load)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or loaded_from_file" 2>&1 | tail -4
echo "== tcgen05 probe (raw accumulators vs layout hypotheses)"
timeout 180 python scripts/tc5_probe.py 2>&1 | tail -12
echo "== tcgen05 parity tests"
B200PIR_TEST_TC5=1 timeout 600 python -m pytest tests/test_gpu_tcgen05.py -x -q 2>&1 | tail -6
echo "== sparse server fold"
B200PIR_TEST_SPARSE_FOLD=1 timeout 300 python -m pytest tests/test_gpu_sparse_fold.py -x -q 2>&1 | tail -3
echo "== 4096-point NTT (config #5)"
B200PIR_TEST_NTT4K=1 timeout 300 python -m pytest tests/test_gpu_ntt4096.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_kernels.py ntt ntt4096 2>&1 | cut -c1-400
echo "== bench: format 1 (8 / 16 per pass) vs format 2"
for args in "--db-format 1 --queries-per-pass 8" "--db-format 1 --queries-per-pass 16" "--db-format 2 --queries-per-pass 16"; do
  timeout 300 python bench.py --no-cpu-baseline $args > gpurun_out/bringup_bench.json 2> gpurun_out/bringup_bench.err
  python - "$args" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bringup_bench.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "->", round(d["value"], 1), "q/s  frac", round(d["roofline"]["frac"], 3), "kernel_ms", round(d["roofline"]["kernel_ms"], 3),
          {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "-> failed:", e, open("gpurun_out/bringup_bench.err").read()[-600:])
PY
done
} 2>&1 | tee gpurun_out/round2_bringup.log
