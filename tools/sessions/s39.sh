#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s39.log 2>&1
echo "=== tests"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_halo_s2_gpu.py tests/test_model_gpu.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c3.py -m gpu -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', round(d['value'],1), round(d['e2e']['value'],1))"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', round(d['value'],1), round(d['e2e']['value'],1))"
