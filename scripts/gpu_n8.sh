#!/bin/bash
# usage: gpu_n8.sh N  -- the driver's multi-GPU bench command (inference line with e2e / e2e_lean / legs) + the training step at N GPUs
N=${1:-8}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 \
   bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"; tail -2 gpurun_out/bench_n$N.err
bash scripts/gpu_train_multi.sh $N
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), "lean", round(d["e2e_lean"]["value"]), round(d["e2e_lean"]["ms_per_step"], 4))
print("legs", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["e2e_legs_ms"].items() if k != "what"})
print("smpl", d.get("smpl_standalone"))
print("train", d.get("training"))
PY
