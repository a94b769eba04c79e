#!/bin/bash
# diagnosis (8 GPUs): per-rank step times of the replica phase of bench --gpus 8 / 4 under a few switches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PORT=30017
run() { np=$1; tag=$2; shift; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $np --steps 5 --warmup 3 --no-configs "$@" > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err; PORT=$((PORT+1)); echo "== $tag"; grep "^rank" gpurun_out/r2j_$tag.err | sort | tr '\n' ';'; echo; python -c "import json;b=json.loads([l for l in open('gpurun_out/r2j_$tag.json') if l.startswith('{')][-1]);print(b['ms_per_step'], b['e2e']['ms_per_step'], b['clocks'])"; }
run 8 n8_default
run 8 n8_nosampler --no-sampler
run 8 n8_nonuma --no-numa-bind
run 8 n8_neither --no-sampler --no-numa-bind
run 4 n4_default
run 4 n4_neither --no-sampler --no-numa-bind
nvidia-smi --query-gpu=index,power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
