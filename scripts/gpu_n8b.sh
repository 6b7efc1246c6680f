#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 50 --warmup 5 --no-hmr --no-fp32 --no-fold --no-live --cpu-budget 0 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "bench n8 exit=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_n8b.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_lean"]["value"], d["e2e_legs_ms"], d["smpl_standalone"]["ms"] if d.get("smpl_standalone") else None, d["training"]["ms_per_step"] if d.get("training") else None)
P
