#!/bin/bash
# round-2 (second half) evidence: ncu launch list of one forward + ncu --set full of every hot kernel at the bench shape.
# Reports -> gpurun_out/r2b_* (summarised into profiles/ by scripts/ncu_summary.py / ncu_counters.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
BENCH="python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold --no-train --no-fp32 --no-hmr"
TP_UM_NOCOOP=1 timeout -s KILL 400 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv $BENCH > gpurun_out/r2b_launches.log 2>&1; echo "launch list exit=$?"
for K in k_gru_umma k_gemm_bf16_tc k_heads_base k_ief_cluster k_smpl_verts_tc; do
  TP_UM_NOCOOP=1 timeout -s KILL 400 $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/r2b_$K $BENCH > gpurun_out/r2b_ncu_$K.log 2>&1; echo "ncu $K exit=$?"
done
timeout -s KILL 400 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench exit=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err; echo "ref exit=$?"
