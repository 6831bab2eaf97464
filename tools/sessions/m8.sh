#!/bin/bash
# 8-GPU runs: replicas C2 (what the driver's scaling run does), crop-sharded C4 / C5 / C2 (north_star: batch sharded over 8 GPUs, one all-gather)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/m8.log 2>&1
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 "$@" 2>&1 | tail -1; }
echo "=== replicas C2"; run --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c2_n8.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n8.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['n_gpus'])"
echo "=== sharded C4 (64 crops = 8 images x 8 persons over 8 GPUs)"; run --steps 10 --warmup 3 --sharded --workload C4 --no-cpu-baseline > gpurun_out/r02_bench_c4_n8_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_n8_sharded.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['config']['partitioning'][:220])"
echo "=== sharded C5"; run --steps 10 --warmup 3 --sharded --workload C5 --no-cpu-baseline > gpurun_out/r02_bench_c5_n8_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_n8_sharded.json')); print(round(d['value'],1), round(d['e2e']['value'],1))"
echo "=== sharded C2"; run --steps 20 --warmup 5 --sharded --no-cpu-baseline > gpurun_out/r02_bench_c2_n8_sharded.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_n8_sharded.json')); print(round(d['value'],1), round(d['e2e']['value'],1))"
