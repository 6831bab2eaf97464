#!/bin/bash
# 2-GPU session: sharded forward tests + sharded bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s8.log 2>&1
nvidia-smi -L
echo "=== sharded tests"; timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -25
echo "=== bench C2 sharded N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --sharded 2>&1 | tail -4
echo "=== bench C2 replicas N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -4
