#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "smpl" > gpurun_out/um_smpl.log 2>&1; echo "smpl tests exit=$? $(tail -1 gpurun_out/um_smpl.log)"
grep -E "^FAILED|^ERROR|Error|assert|^E " gpurun_out/um_smpl.log | head -12
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 random
timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 coherent
TP_SMPL_UM=1 timeout 120 python scripts/smpl_standalone.py 65536 bf16 5 random
