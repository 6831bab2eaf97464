#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s16.log 2>&1
echo "=== sanity (deeper A ring)"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "=== fc1 80->320 split gelu"; timeout 200 python tools/trace_halo_problem.py 80 320 1 1 1 0 8 64 48 6 2>&1 | tail -30
echo "=== fc2 320->80 split gelu+res"; timeout 200 python tools/trace_halo_problem.py 320 80 1 1 1 1 8 64 48 6 2>&1 | tail -30
echo "=== q 80->96 split"; timeout 200 python tools/trace_halo_problem.py 80 96 1 1 0 0 8 70 49 6 2>&1 | tail -30
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
echo "=== bench C2"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'])"
echo "=== bench C3"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"
