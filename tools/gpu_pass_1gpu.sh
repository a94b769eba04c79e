#!/bin/bash
# round-2 evidence pass on one B200: full -m gpu suite with the radix-sort densify + MSM v2, bench, launch lists, ncu --set full
# captures of the new kernels, sanitizers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1
echo "t1 rc=$?"; tail -6 gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], b['throughput_batched'], b['config']['no_tables'])
    for r in b['configs']: print({k:r[k] for k in ('name','densify_ms','commit_ms','prove_ms','golden_match')})
except Exception as e: print('no bench line', e)
PY
timeout 600 python bench.py --workload msm --msm-max-log 24 > gpurun_out/r2_msm.json 2> gpurun_out/r2_msm.err
echo "msm rc=$?"; python - <<PY
import json
try:
    m=json.loads([l for l in open('gpurun_out/r2_msm.json') if l.startswith('{')][-1])
    for r in m['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('log_n','scalars','ms','c','windows','frac_of_add_ceiling','same_point_as_cpu')})
except Exception as e: print('no msm line', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_launches_prove.csv python tools/prove_once.py 20 2 > gpurun_out/r2_prove_once.log 2>&1
echo "ncu prove rc=$?"; tail -2 gpurun_out/r2_prove_once.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_msm.csv python tools/msm_once.py 16 22 > gpurun_out/r2_msm_once.log 2>&1
echo "ncu msm rc=$?"; tail -3 gpurun_out/r2_msm_once.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accum_kernel -s 1 -c 1 -o gpurun_out/r2_prof_msm_accum python tools/msm_once.py 22 > gpurun_out/r2_ncu1.log 2>&1
echo "ncu accum rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bullet_fused_kernel -s 50 -c 1 -o gpurun_out/r2_prof_bullet_fused python tools/prove_once.py 20 2 > gpurun_out/r2_ncu2.log 2>&1
echo "ncu bullet rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dz_radix_scatter_kernel -s 2 -c 1 -o gpurun_out/r2_prof_radix_scatter python tools/prove_once.py 20 2 > gpurun_out/r2_ncu3.log 2>&1
echo "ncu radix rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_synccheck.log 2>&1
echo "synccheck rc=$?"; tail -4 gpurun_out/r2_synccheck.log
LASSO_SHARD_SAME_GPU=1 timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29719 tools/sharded_check.py 2 4 16 0 4096 1 > gpurun_out/r2_memcheck_sharded.log 2>&1
echo "memcheck sharded rc=$?"; grep -E "ERROR SUMMARY|SHARDED_CHECK|case" gpurun_out/r2_memcheck_sharded.log | tail -6
