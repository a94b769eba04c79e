#!/bin/bash
# round-2 evidence pass on 8 x B200: the driver's scaling runs at HEAD — bench --gpus 8 and --gpus 4 (replicas + sharded configs),
# MSM sweep on 8 GPUs; host CPU limits of the box for the e2e analysis
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; lscpu | egrep "Model name|Socket|Core|Thread|NUMA" ) > gpurun_out/r2_cpu.txt 2>&1
cat gpurun_out/r2_cpu.txt
PORT=29817
for NP in 8 4; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((PORT+NP)) bench.py --gpus $NP --steps 3 --warmup 3 > gpurun_out/r2_bench_n$NP.json 2> gpurun_out/r2_bench_n$NP.err
echo "bench N=$NP rc=$?"; tail -2 gpurun_out/r2_bench_n$NP.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2_bench_n$NP.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, b['e2e'], b['config'].get('numa_node'))
    for r in b['configs']: print({k:r.get(k) for k in ('name','densify_ms','commit_ms','prove_ms','ms_per_proof','single_gpu_ms_per_proof','matches_single_gpu','golden_match')})
except Exception as e: print('no bench line', e)
PY
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((PORT+20)) bench.py --gpus 8 --workload msm --msm-max-log 26 > gpurun_out/r2_msm_n8.json 2> gpurun_out/r2_msm_n8.err
echo "msm N=8 rc=$?"; python - <<PY
import json
try:
    m=json.loads([l for l in open('gpurun_out/r2_msm_n8.json') if l.startswith('{')][-1])
    for r in m['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('log_n','scalars','ms','c','windows','same_point_as_cpu')})
except Exception as e: print('no msm line', e)
PY
