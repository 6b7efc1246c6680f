#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "smpl" > gpurun_out/um_smpl.log 2>&1; echo "smpl tests exit=$? $(tail -1 gpurun_out/um_smpl.log)"
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 random
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 coherent
for skin in random coherent; do
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_lbs_um2 -s 2 -c 1 -f -o gpurun_out/prof_k_smpl_lbs_um2_$skin \
   python scripts/smpl_standalone.py 16384 bf16 1 $skin > gpurun_out/ncu_k_smpl_lbs_um2_$skin.log 2>&1; echo "ncu exit=$?"
done
