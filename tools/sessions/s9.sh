#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s9.log 2>&1
echo "=== mma2 probe"; timeout 120 ./build/mma2_probe.bin 2>&1 | tail -12
echo "=== postproc tests"; timeout 600 python -m pytest tests/test_postproc.py -m gpu -q 2>&1 | tail -15
echo "=== bench C2"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 2>&1 | tail -2
echo "=== bench C5"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C5 2>&1 | tail -2
echo "=== bench C3"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C3 2>&1 | tail -2
