#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s11.log 2>&1
echo "=== stress (pair auto)"; timeout 600 python -m pytest tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -30
echo "=== kernels, pair forced"; I2R_HALO_PAIR=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -30
echo "=== kernels, pair auto"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -8
echo "=== models"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_gpu_c3.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c45.py -m gpu -q 2>&1 | tail -12
echo "=== bench C2"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-1200
echo "=== bench C2 pair off"; I2R_HALO_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-400
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 2>&1 | tail -2 | cut -c1-400
echo "=== bench C4 serial branches"; I2R_CONCURRENT_BRANCHES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 2>&1 | tail -2 | cut -c1-400
echo "=== bench C3"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C3 2>&1 | tail -2 | cut -c1-400
