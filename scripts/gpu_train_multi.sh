#!/bin/bash
# usage: gpu_train_multi.sh N   -- training step on N GPUs of one box (torchrun, NCCL)
N=${1:-2}
mkdir -p gpurun_out
for cfg in "fp32:" "bf16:--tc-grads"; do
  prec=${cfg%%:*}; extra=${cfg#*:}
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --config train --steps 10 --warmup 3 --train-precision $prec $extra > gpurun_out/bench_train_${prec}_n$N.json 2> gpurun_out/bench_train_${prec}_n$N.err
  echo "train $prec N=$N exit=$?"; tail -2 gpurun_out/bench_train_${prec}_n$N.err
done
python - <<PY
import json
for prec in ("fp32", "bf16"):
    f = f"gpurun_out/bench_train_{prec}_n$N.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        t = d["training"]
        print(f, "seq/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "no-allreduce ms", t.get("ms_per_step_without_allreduce"),
              "exposed", t.get("exposed_comm_frac"), "bytes", t.get("allreduce_bytes_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
