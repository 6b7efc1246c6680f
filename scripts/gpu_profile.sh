#!/bin/bash
# ncu evidence for profiles/: launch list of one forward + --set full captures of the top kernels.
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
ARGS="bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python $ARGS > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit=$?"
python scripts/launch_table.py gpurun_out/launches.csv > gpurun_out/launch_table.txt; tail -25 gpurun_out/launch_table.txt
for k in k_gru_bf16_tma k_gemm_bf16_tc k_skinny_bf16 k_smpl_verts; do
  timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$k python $ARGS > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit=$?"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after.txt
