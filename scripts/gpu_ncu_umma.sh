#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
TP_UM_NOCOOP=1 timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_gru_umma -s 3 -c 1 -f -o gpurun_out/prof_k_gru_umma \
   python bench.py --steps 1 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold > gpurun_out/ncu_k_gru_umma.log 2>&1; echo "ncu exit=$?"
tail -5 gpurun_out/ncu_k_gru_umma.log
