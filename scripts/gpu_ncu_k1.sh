#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
BENCH="python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold --no-train --no-fp32 --no-hmr"
TP_UM_NOCOOP=1 timeout -s KILL 300 $NCU --set full --clock-control none --import-source on -k regex:k_gemm_bf16_tc -s 3 -c 1 -f -o gpurun_out/r2b_k_gemm_bf16_tc $BENCH > gpurun_out/r2b_ncu_k1.log 2>&1; echo "ncu exit=$?"
ls -la gpurun_out/r2b_k_gemm_bf16_tc.ncu-rep
