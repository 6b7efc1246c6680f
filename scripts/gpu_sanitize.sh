#!/bin/bash
# compute-sanitizer over the GPU parity tests of the persistent / hand-synchronised kernels:
#   memcheck  : out-of-bounds / misaligned global + shared accesses
#   racecheck : shared-memory hazards (the warp-specialised kernels hand buffers over through mbarriers / named barriers)
# K2 = GRU recurrence kernels, K3 = heads + IEF, K5 = SMPL kernels (small-batch fused, large-batch tcgen05).  Logs -> gpurun_out/.
mkdir -p gpurun_out
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
SEL='test_smpl_large_batch_h36m or test_gru_recurrence_umma or test_gru_recurrence_two_interleaved_directions or test_smpl_forward_tensor_core_blend or test_smpl_large_batch_split_path or test_smpl_forward_all_pose_kinds'
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 30 --error-exitcode 9 \
     python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitizer_${tool}_kernels.log 2>&1
  echo "$tool kernels exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_kernels.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_${tool}_kernels.log | tail -1)"
  timeout 1500 $CS --tool $tool --print-limit 30 --error-exitcode 9 \
     python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -x -k "golden and bf16 and (H64_B2_T4 or H128)" > gpurun_out/sanitizer_${tool}_e2e.log 2>&1
  echo "$tool e2e exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_e2e.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_${tool}_e2e.log | tail -1)"
done
