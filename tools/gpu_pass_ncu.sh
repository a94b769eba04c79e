#!/bin/bash
# round-2 ncu captures (one B200): ncu --set full of the bind kernel (new shape), the big fused grand-product round and the LT
# evaluation kernel on the LT C=8 2^22 configuration
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bind_top2_kernel -s 3 -c 1 -o gpurun_out/r2_prof_bind_top2 python tools/bind_only.py > gpurun_out/r2_ncu0.log 2>&1
echo "ncu bind rc=$?"; tail -2 gpurun_out/r2_ncu0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_bind_eval_cubic_comb_kernel -s 0 -c 1 -o gpurun_out/r2_prof_bindeval_lt python tools/spans_config.py lt_c8_s22 1 > gpurun_out/r2_ncu1.log 2>&1
echo "ncu bindeval rc=$?"; tail -2 gpurun_out/r2_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_eval_lt_kernel -s 0 -c 1 -o gpurun_out/r2_prof_evallt python tools/spans_config.py lt_c8_s22 1 > gpurun_out/r2_ncu2.log 2>&1
echo "ncu evallt rc=$?"; tail -2 gpurun_out/r2_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_eval_cubic_comb_kernel -s 0 -c 1 -o gpurun_out/r2_prof_evalcubic_lt python tools/spans_config.py lt_c8_s22 1 > gpurun_out/r2_ncu3.log 2>&1
echo "ncu evalcubic rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-batched 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(b['ms_per_step'], b['e2e']['ms_per_step'], b['roofline'])"
