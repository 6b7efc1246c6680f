#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_lbs_um2 -s 2 -c 1 -f -o gpurun_out/prof_k_smpl_lbs_um2 \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/ncu_k_smpl_lbs_um2.log 2>&1; echo "ncu exit=$?"
