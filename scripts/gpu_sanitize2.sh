#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests of the kernels written in the second half of round 2:
# k_heads_base + k_ief_cluster (bulk copies, mbarriers, st.async into distributed shared memory), the persistent tcgen05 GEMM,
# the cta_group::2 pair GEMM (child process of test_gemm_bf16_tcgen05_cta_pairs), the HMR data-movement kernels.  Logs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
SEL='test_ief_cluster_kernel or (test_gemm_bf16_tcgen05 and not cta_pairs and (8-192-128 or 512-768-2176 or 300-9216-64 or 512-1152-192))'
timeout -s KILL 900 $CS --tool memcheck --print-limit 30 --error-exitcode 9 \
   python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitizer2_memcheck_kernels.log 2>&1
echo "memcheck kernels exit=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer2_memcheck_kernels.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_memcheck_kernels.log | tail -1)"
TP_TC_2CTA=1 timeout -s KILL 600 $CS --tool memcheck --print-limit 30 --error-exitcode 9 \
   python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "test_gemm_bf16_tcgen05 and not cta_pairs and (300-9216-64 or 512-9216-256)" > gpurun_out/sanitizer2_memcheck_2cta.log 2>&1
echo "memcheck 2cta exit=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer2_memcheck_2cta.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_memcheck_2cta.log | tail -1)"
timeout -s KILL 900 $CS --tool memcheck --print-limit 30 --error-exitcode 9 \
   python -m pytest tests/test_gpu_e2e.py tests/test_hmr.py -m gpu -q --no-header -p no:cacheprovider -x -k "(golden and bf16 and H2048) or fused_heads or conv_data_movement or gemm_epilogue" > gpurun_out/sanitizer2_memcheck_e2e.log 2>&1
echo "memcheck e2e exit=$? $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer2_memcheck_e2e.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_memcheck_e2e.log | tail -1)"
timeout -s KILL 900 $CS --tool racecheck --print-limit 30 --error-exitcode 9 \
   python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "test_ief_cluster_kernel" > gpurun_out/sanitizer2_racecheck_kernels.log 2>&1
echo "racecheck kernels exit=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer2_racecheck_kernels.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_racecheck_kernels.log | tail -1)"
timeout -s KILL 900 $CS --tool racecheck --print-limit 30 --error-exitcode 9 \
   python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -x -k "fused_heads" > gpurun_out/sanitizer2_racecheck_e2e.log 2>&1
echo "racecheck e2e exit=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer2_racecheck_e2e.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_racecheck_e2e.log | tail -1)"
