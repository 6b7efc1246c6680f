#!/bin/bash
cd "$(dirname "$0")/.."
timeout -s KILL 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "non_current_device" 2>&1 | tail -2
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-hmr > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_lean"]["value"], d["training"]["ms_per_step"] if d.get("training") else None)
P
tail -3 gpurun_out/bench_n2.err
