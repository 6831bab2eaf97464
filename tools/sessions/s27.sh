#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s27.log 2>&1
echo "=== kernel + model tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_halo_s2_gpu.py tests/test_halo_stress_gpu.py tests/test_model_gpu.py tests/test_hrformer_kernels_gpu.py tests/test_model_gpu_hrt.py -m gpu -q -x 2>&1 | tail -6
echo "=== layer1 conv3: 64->256 1x1 + residual, 32 crops"; timeout 200 python tools/trace_halo_problem.py 64 256 1 0 0 1 32 64 48 48 2>&1 | tail -5
echo "=== ds: 64->256 1x1, no residual"; timeout 200 python tools/trace_halo_problem.py 64 256 1 0 0 0 32 64 48 48 2>&1 | tail -5
echo "=== HRT fc1 80->320 split gelu 8 crops"; timeout 200 python tools/trace_halo_problem.py 80 320 1 1 1 0 8 64 48 48 2>&1 | tail -5
for wl in C2 C3 C4 C5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
