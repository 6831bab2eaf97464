#!/bin/bash
# GPU session 4: halo stress/determinism, engine tests, full suite without the crop-grouping workaround, quick bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s7.log 2>&1
echo "=== stress"; timeout 900 python -m pytest tests/test_halo_stress_gpu.py -m gpu -q 2>&1 | tail -25
echo "=== engine"; timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q 2>&1 | tail -25
echo "=== suite"; timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_halo_stress_gpu.py --deselect tests/test_engine_gpu.py 2>&1 | tail -25
echo "=== bench C2"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -3
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 2>&1 | tail -3
