#!/bin/bash
# HEAD check of the hybrid primary sumcheck (fused rounds from q = 2^15 up): end-to-end parity + golden at size, then timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_big_configs.py tests/test_gpu_prove.py tests/test_gpu_kernels.py -m gpu -x -q -k "big or prove or bind_round" > gpurun_out/final_t.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/final_t.log
timeout 40 python tools/ab_primary.py 20 12 > gpurun_out/ab_s20_hybrid.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/ab_s20_hybrid.log
