#!/bin/bash
# launch list of one forward + full ncu capture of the dominant kernel (k_gru_bf16_dual)
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
python scripts/launch_table.py gpurun_out/launches.csv | tee gpurun_out/launch_table.txt
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_gru_bf16_dual -s 3 -c 1 -f -o gpurun_out/prof_k_gru_bf16_dual \
   python bench.py --steps 1 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold > gpurun_out/ncu_dual.log 2>&1; echo "ncu dual exit=$?"
