#!/bin/bash
# Profiling session: butterfly-rate microbenchmark, ncu captures of the fold / expansion kernels, launch list of a bench step.
TAG=${1:-r02}
mkdir -p gpurun_out
{
echo "== tests touched since the last full run"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_bench_config.py tests/test_gpu_parity.py -x -q -k "tc5_long or coalesced or dpir or config4 or config5 or error_behaviour" 2>&1 | tail -12
echo "== butterfly-rate microbenchmark"
timeout 120 scripts/ubench/bfly
echo "== ncu: fold kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fold_res_lz -s 4 -c 1 -o gpurun_out/ncu_fold_${TAG} -f \
  python bench.py --steps-only --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_fold_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_fold_${TAG}.log | cut -c1-200
echo "== ncu: expansion kernel (widest round)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_expand_round_res -s 15 -c 1 -o gpurun_out/ncu_expand_${TAG} -f \
  python bench.py --steps-only --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_expand_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_expand_${TAG}.log | cut -c1-200
echo "== launch list of the timed steps"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps-only --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/launches_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.log | cut -c1-200
} 2>&1 | tee gpurun_out/gpu_prof_${TAG}.log
