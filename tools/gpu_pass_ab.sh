#!/bin/bash
# fused primary-sumcheck round: resident CTAs per SM (2 or 3) x fused/unfused, spans on, 2^24 then 2^20 lookups
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mb in 2 3; do
  LASSO_B200_FUSED_MINB=$mb LASSO_B200_SPANS=1 timeout 200 python tools/ab_primary.py 24 6 > gpurun_out/ab_s24_minb$mb.log 2>&1; echo "minb=$mb s24 rc=$?"; tail -3 gpurun_out/ab_s24_minb$mb.log
done
for mb in 2 3; do
  LASSO_B200_FUSED_MINB=$mb LASSO_B200_SPANS=1 timeout 200 python tools/ab_primary.py 20 16 > gpurun_out/ab_s20_minb$mb.log 2>&1; echo "minb=$mb s20 rc=$?"; tail -4 gpurun_out/ab_s20_minb$mb.log
done
