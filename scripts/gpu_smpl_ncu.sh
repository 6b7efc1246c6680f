#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_skin -s 40 -c 1 -f -o gpurun_out/prof_smpl_skin \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/ncu_smpl_skin.log 2>&1; echo "ncu exit=$?"
