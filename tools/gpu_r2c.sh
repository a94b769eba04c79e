#!/bin/bash
# round-2 GPU pass C (8 GPUs): sharded parity on 8 ranks, bench --gpus 8 (replicas + the three configs sharded), MSM sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NP=${1:-8}
nvidia-smi -L > gpurun_out/r2c_gpus.txt; nvidia-smi topo -m >> gpurun_out/r2c_gpus.txt 2>&1
PORT=29617
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT tools/sharded_check.py > gpurun_out/r2c_sharded_check.log 2>&1
echo "sharded_check N=$NP rc=$?"; grep -E "case|collective|SHARDED" gpurun_out/r2c_sharded_check.log | tail -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $NP --steps 3 --warmup 3 > gpurun_out/r2c_bench_n$NP.json 2> gpurun_out/r2c_bench_n$NP.err
echo "bench N=$NP rc=$?"; tail -3 gpurun_out/r2c_bench_n$NP.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2c_bench_n$NP.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, b['e2e'], b['config'].get('numa_node'))
    for r in b['configs']: print(r)
except Exception as e: print('no bench line', e)
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((PORT+2)) bench.py --gpus $NP --workload msm --msm-max-log 26 > gpurun_out/r2c_msm_n$NP.json 2> gpurun_out/r2c_msm_n$NP.err
echo "msm N=$NP rc=$?"; tail -3 gpurun_out/r2c_msm_n$NP.err; python - <<PY
import json
try:
    m=json.loads([l for l in open('gpurun_out/r2c_msm_n$NP.json') if l.startswith('{')][-1])
    for r in m['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e: print('no msm line', e)
PY
