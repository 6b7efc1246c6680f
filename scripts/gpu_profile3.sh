#!/bin/bash
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
for K in k_gemm_bf16_tc k_ief_fused k_smpl_verts_tc; do
  timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$K \
     python bench.py --steps 1 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold > gpurun_out/ncu_$K.log 2>&1; echo "ncu $K exit=$?"
done
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_skin -s 40 -c 1 -f -o gpurun_out/prof_k_smpl_skin \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/ncu_skin.log 2>&1; echo "ncu skin exit=$?"
