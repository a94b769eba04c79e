#!/bin/bash
# diagnosis (2 GPUs): why the replica steps of bench --gpus N got slower than a lone proof on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PORT=29917
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 5 --warmup 3 --no-configs "$@" > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; PORT=$((PORT+1)); echo "== $tag"; grep "^rank" gpurun_out/r2i_$tag.err; python -c "import json;b=json.loads([l for l in open('gpurun_out/r2i_$tag.json') if l.startswith('{')][-1]);print(b['ms_per_step'], b['e2e']['ms_per_step'])"; }
run default
run nosampler --no-sampler
run nonuma --no-numa-bind
run neither --no-sampler --no-numa-bind
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-batched > gpurun_out/r2i_n1.json 2> gpurun_out/r2i_n1.err; grep "^rank" gpurun_out/r2i_n1.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-batched --no-sampler > gpurun_out/r2i_n1ns.json 2> gpurun_out/r2i_n1ns.err; grep "^rank" gpurun_out/r2i_n1ns.err
