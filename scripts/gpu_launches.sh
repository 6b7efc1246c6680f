#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
python scripts/launch_table.py gpurun_out/launches.csv
