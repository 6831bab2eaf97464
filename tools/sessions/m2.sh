#!/bin/bash
# 2-GPU validation: sharded-forward tests (NCCL), replicas bench, crop-sharded bench (C2 and C4 with 2 images)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/m2.log 2>&1
echo "=== 2-GPU tests"; timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -4
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@" 2>&1 | tail -1; }
echo "=== replicas C2"; run --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n2.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['n_gpus'])"
echo "=== sharded C2"; run --steps 20 --warmup 5 --sharded > gpurun_out/r02_bench_c2_n2_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n2_sharded.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['config']['partitioning'][:200])"
echo "=== sharded C4"; run --steps 10 --warmup 3 --sharded --workload C4 > gpurun_out/r02_bench_c4_n2_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_n2_sharded.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['config']['partitioning'][:200])"
