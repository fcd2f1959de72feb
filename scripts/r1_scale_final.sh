#!/bin/bash
# final multi-GPU numbers: S8 default at N GPUs, and (N = 8) BASELINE config #3 (S256, 128 concurrent queries)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scalef_n$N.json 2> gpurun_out/scalef_n$N.err
echo "rc=$?"
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --workload S256 --steps 3 --warmup 3 > gpurun_out/scalef_s256.json 2> gpurun_out/scalef_s256.err
  echo "rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/scalef_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["config"]["batch"], d["roofline"]["frac"], d["stage_ms_per_step"])
    except Exception as e:
        print(f, "parse failed", e)
PY
