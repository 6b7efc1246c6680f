#!/bin/bash
cd "$(dirname "$0")/.."
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout -s KILL 300 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_lbs_um2 -s 2 -c 1 -f -o gpurun_out/r2b_k_smpl_lbs_um2 \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/r2b_ncu_k_smpl_lbs_um2.log 2>&1; echo "ncu um2 exit=$?"
timeout -s KILL 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_smpl65536_launches.csv python scripts/smpl_standalone.py 65536 bf16 1 > gpurun_out/r2b_smpl_launches.log 2>&1; echo "smpl launch list exit=$?"
