#!/bin/bash
cd "$(dirname "$0")/.."
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout -s KILL 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/hmr_launches.csv python scripts/hmr_time.py > gpurun_out/hmr_launches.log 2>&1; echo "exit=$?"
