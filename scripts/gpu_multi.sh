#!/bin/bash
# Multi-GPU session: bash scripts/gpu_multi.sh N [tag]   (under gpurun --gpus N)
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
run() {
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/multi.json 2> gpurun_out/multi.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/multi.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("%-28s %8.1f q/s  e2e %8.1f  ms/step %.3f  verified %s " % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("verified")),
          {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "failed:", e, open("gpurun_out/multi.err").read()[-1500:])
PY
}
{
echo "== $N GPUs"
run "copy engines, pipelined (default)" X=1 -- --timeline
grep -A34 "rank 0 timeline" gpurun_out/multi.err | head -36
cp gpurun_out/multi.json gpurun_out/bench_${TAG}_n${N}.json
tail -3 gpurun_out/multi.err | cut -c1-300
if [ -n "$TIMELINE" ]; then
run "default + timeline" X=1 -- --timeline --no-verify --steps-only
grep -A200 "rank 0 timeline" gpurun_out/multi.err | head -120
run "8 hardware queues" CUDA_DEVICE_MAX_CONNECTIONS=8 -- --no-verify --steps-only
run "no pipeline, 2 waves" X=1 -- --no-pipeline --no-verify --steps-only
run "pipeline, 2 waves" X=1 -- --waves 2 --no-verify --steps-only
fi
if [ "$N" = "8" ] && [ -z "$TIMELINE" ] && [ -z "$SKIP_S256" ]; then
echo "== BASELINE configs[2]: S256 (32 GiB plaintext = 256 GiB packed) over 8 GPUs, 128 concurrent queries"
run "S256, 128 queries" X=1 -- --workload S256 --steps 5
cp gpurun_out/multi.json gpurun_out/bench_${TAG}_s256_n${N}.json
tail -3 gpurun_out/multi.err | cut -c1-300
fi
if [ "$N" = "2" ]; then
run "copy engines, 2 waves, no pipeline" X=1 -- --no-pipeline
run "nccl all-gather, 2 waves" X=1 -- --exchange nccl
fi
} 2>&1 | tee gpurun_out/gpu_multi_${TAG}_n${N}.log
