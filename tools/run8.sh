#!/bin/bash
# multi-GPU evidence run: parity of the sharded proof, replica scaling, sharded big configs
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29601 tools/sharded_check.py 2 4 16 0 4096 1 2>&1 | grep -E "case|SHARDED_CHECK"
timeout 400 $TR --master-port 29602 tools/sharded_check.py 3 8 8 0 2048 0 2>&1 | grep -E "case|SHARDED_CHECK"
timeout 400 $TR --master-port 29603 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-330
timeout 400 $TR --master-port 29604 bench.py --gpus $N --steps 3 --warmup 3 --sharded 2>&1 | tail -1 | cut -c1-330
SHARDED=1 timeout 600 $TR --master-port 29605 tools/big_config.py 3 8 16 0 22 nocheck 3 2>&1 | grep -E "kind="
SHARDED=1 timeout 600 $TR --master-port 29606 tools/big_config.py 4 4 16 40 24 nocheck 3 2>&1 | grep -E "kind="
