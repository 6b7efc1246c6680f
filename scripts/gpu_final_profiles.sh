#!/bin/bash
# final-build evidence: ncu launch list of one forward + ncu --set full of every hot kernel at the bench shape, and of the large-batch
# SMPL kernel at 16,384 bodies.  Reports -> gpurun_out/ (summarised into profiles/ by scripts/ncu_summary.py / ncu_counters.py).
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
BENCH="python bench.py --steps 2 --warmup 3 --cpu-budget 0 --no-graph --no-live --no-smpl --no-fold --no-train --no-fp32"
TP_UM_NOCOOP=1 timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv $BENCH > gpurun_out/final_launches.log 2>&1; echo "launch list exit=$?"
for K in k_gru_umma k_gemm_bf16_tc k_ief_fused k_smpl_verts_tc k_pack_rows; do
  TP_UM_NOCOOP=1 timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/final_$K $BENCH > gpurun_out/final_ncu_$K.log 2>&1; echo "ncu $K exit=$?"
done
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_smpl_lbs_um2 -s 2 -c 1 -f -o gpurun_out/final_k_smpl_lbs_um2 \
   python scripts/smpl_standalone.py 16384 bf16 1 > gpurun_out/final_ncu_k_smpl_lbs_um2.log 2>&1; echo "ncu um2 exit=$?"
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_smpl65536_launches.csv python scripts/smpl_standalone.py 65536 bf16 1 > gpurun_out/final_smpl_launches.log 2>&1; echo "smpl launch list exit=$?"
