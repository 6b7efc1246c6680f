#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU) + reference arm
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 30 --warmup 5 --no-live > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"
tail -n 1 gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],4),'smpl',d.get('smpl_standalone'))"
tail -3 gpurun_out/bench_n$N.err
