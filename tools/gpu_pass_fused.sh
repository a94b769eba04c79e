#!/bin/bash
# primary sumcheck: bind fused into the next round's evaluation + shift-based weights — parity suite, before/after timing,
# launch list and one ncu --set full capture of the fused kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f_t1.log 2>&1
echo "t1 rc=$?"; tail -5 gpurun_out/f_t1.log
for mode in fused unfused; do
  if [ $mode = unfused ]; then export LASSO_B200_UNFUSED_PRIMARY=1; else unset LASSO_B200_UNFUSED_PRIMARY; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-configs --no-batched --no-cpu-baseline > gpurun_out/f_bench_$mode.json 2> gpurun_out/f_bench_$mode.err
  echo "bench $mode rc=$?"; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/f_bench_$mode.json') if l.startswith('{')][-1])
    print('$mode', b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches'], b['config'].get('golden_match'), b['roofline']['frac'])
except Exception as e: print('no bench line', e)
PY
done
unset LASSO_B200_UNFUSED_PRIMARY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/f_launches_prove.csv python tools/prove_once.py 20 2 > gpurun_out/f_prove_once.log 2>&1
echo "ncu prove rc=$?"; tail -2 gpurun_out/f_prove_once.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sc_bind_eval_linear_kernel -s 19 -c 1 -o gpurun_out/f_prof_bind_eval_linear python tools/prove_once.py 20 2 > gpurun_out/f_ncu1.log 2>&1
echo "ncu fused rc=$?"; tail -2 gpurun_out/f_ncu1.log
