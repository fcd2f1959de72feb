#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
for w in 2 1 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --waves $w > gpurun_out/scale${N}_w$w.json 2> gpurun_out/scale${N}_w$w.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale${N}_w$w.json").read().strip().splitlines()[-1])
    print("N $N waves $w", d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["batch"], d["stage_ms_per_step"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/scale${N}_w$w.err").read()[-2000:])
PY
done
