#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> [skip]
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
K=$1; SKIP=${2:-8}
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_$K \
   python bench.py --steps 1 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl > gpurun_out/ncu_$K.log 2>&1; echo "ncu $K exit=$?"
