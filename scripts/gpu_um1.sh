#!/bin/bash
# fused tcgen05 blend + skinning kernel: parity tests, then timing at 65,536 bodies against the GEMM + skin pair
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "smpl" > gpurun_out/um_smpl.log 2>&1; echo "smpl tests exit=$? $(tail -1 gpurun_out/um_smpl.log)"
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/um_smpl.log | head
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5
TP_SMPL_UM=0 timeout 120 python scripts/smpl_standalone.py 65536 bf16 5
timeout 120 python scripts/smpl_standalone.py 4096 bf16 20
