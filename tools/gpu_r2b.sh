#!/bin/bash
# round-2 GPU pass B (2 GPUs): fused opening rounds + MSM skew fix on one GPU, then the sharded proof / MSM on two
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NP=${1:-2}
nvidia-smi -L > gpurun_out/r2b_gpus.txt; nvidia-smi topo -m >> gpurun_out/r2b_gpus.txt 2>&1
python -m pytest tests/test_gpu_prove.py tests/test_golden.py tests/test_gpu_msm_large.py tests/test_gpu_big_configs.py -m gpu -x -q > gpurun_out/r2b_t1.log 2>&1
echo "t1 rc=$?"; tail -4 gpurun_out/r2b_t1.log
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r2b_t2.log 2>&1
echo "t2 rc=$?"; tail -12 gpurun_out/r2b_t2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err
echo "bench1 rc=$?"; tail -2 gpurun_out/r2b_bench1.err; cut -c1-400 gpurun_out/r2b_bench1.json
LASSO_B200_UNFUSED_ROUNDS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2b_bench1_unfused.json 2> gpurun_out/r2b_bench1u.err
echo "bench1 unfused rc=$?"; cut -c1-300 gpurun_out/r2b_bench1_unfused.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2b_launches_prove.csv python tools/prove_once.py 20 2 > gpurun_out/r2b_prove_once.log 2>&1
echo "ncu prove rc=$?"; tail -2 gpurun_out/r2b_prove_once.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2b_launches_msm.csv python tools/msm_once.py 16 18 22 > gpurun_out/r2b_msm_once.log 2>&1
echo "ncu msm rc=$?"; tail -4 gpurun_out/r2b_msm_once.log
PORT=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NP --steps 3 --warmup 3 > gpurun_out/r2b_bench_n$NP.json 2> gpurun_out/r2b_bench_n$NP.err
echo "bench N=$NP rc=$?"; tail -3 gpurun_out/r2b_bench_n$NP.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2b_bench_n$NP.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, b['e2e'])
    for r in b['configs']: print(r)
except Exception as e: print('no bench line', e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $NP --workload msm --msm-max-log 24 > gpurun_out/r2b_msm_n$NP.json 2> gpurun_out/r2b_msm_n$NP.err
echo "msm N=$NP rc=$?"; tail -3 gpurun_out/r2b_msm_n$NP.err; cut -c1-600 gpurun_out/r2b_msm_n$NP.json
timeout 600 python bench.py --workload msm --msm-max-log 24 > gpurun_out/r2b_msm_n1.json 2> gpurun_out/r2b_msm_n1.err
echo "msm N=1 rc=$?"; cut -c1-300 gpurun_out/r2b_msm_n1.json
