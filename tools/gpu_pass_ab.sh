#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/ab_primary.py 20 40 > gpurun_out/ab_primary.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/ab_primary.log
LASSO_B200_SPANS=1 timeout 200 python tools/ab_primary.py 20 20 > gpurun_out/ab_primary_spans.log 2>&1; echo "ab spans rc=$?"; tail -4 gpurun_out/ab_primary_spans.log
