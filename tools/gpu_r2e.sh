#!/bin/bash
# round-2 GPU pass E (1 GPU): pipelined densify upload, inlined MSM final kernel; spans of the three configurations
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_prove.py tests/test_gpu_msm_large.py tests/test_gpu_big_configs.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2e_t1.log 2>&1
echo "t1 rc=$?"; tail -4 gpurun_out/r2e_t1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2e_bench.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2e_bench.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], b['throughput_batched'])
except Exception as e: print('no bench line', e)
PY
for K in 2 3 6 8; do timeout 300 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --batch $K 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K=$K', b['throughput_batched'])"; done
timeout 600 python bench.py --workload msm --msm-max-log 24 > gpurun_out/r2e_msm.json 2> gpurun_out/r2e_msm.err
echo "msm rc=$?"; python - <<PY
import json
try:
    m=json.loads([l for l in open('gpurun_out/r2e_msm.json') if l.startswith('{')][-1])
    for r in m['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('log_n','scalars','ms','c','windows','same_point_as_cpu')})
except Exception as e: print('no msm line', e)
PY
for cfg in xor_c4_s20 lt_c8_s22 rc40_c4_s24; do timeout 600 python tools/spans_config.py $cfg 2 > gpurun_out/r2e_spans_$cfg.log 2>&1; tail -2 gpurun_out/r2e_spans_$cfg.log; done
