#!/bin/bash
# 2 GPUs: dedicated-core pinning check + the LT two-lane kernel (tests + spans)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_prove.py tests/test_gpu_big_configs.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2k_t1.log 2>&1
echo "t1 rc=$?"; tail -3 gpurun_out/r2k_t1.log
timeout 600 python tools/spans_config.py lt_c8_s22 2 > gpurun_out/r2k_spans_lt.log 2>&1; tail -2 gpurun_out/r2k_spans_lt.log
PORT=30117
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT+i)) bench.py --gpus 2 --steps 5 --warmup 3 --no-configs > gpurun_out/r2k_b$i.json 2> gpurun_out/r2k_b$i.err
grep "^rank" gpurun_out/r2k_b$i.err | sort
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT+5)) bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_full.json 2> gpurun_out/r2k_full.err
grep "^rank" gpurun_out/r2k_full.err | sort; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2k_full.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, b['e2e'])
    for r in b['configs']: print({k:r.get(k) for k in ('name','densify_ms','commit_ms','prove_ms','ms_per_proof','single_gpu_ms_per_proof','matches_single_gpu','golden_match')})
except Exception as e: print('no bench line', e)
PY
