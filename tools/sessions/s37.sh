#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s37.log 2>&1
export I2R_HALO_SPLIT_ONCE=2
echo "=== tests (stage once everywhere)"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_hrformer_kernels_gpu.py tests/test_model_gpu_c3.py tests/test_model_gpu_hrt.py -m gpu -q -x 2>&1 | tail -4
echo "=== 192->192 3x3 split residual 16 crops 16x12"; timeout 200 python tools/trace_halo_problem.py 192 192 9 1 0 1 16 16 12 16 2>&1 | tail -1
echo "=== 96->96 3x3 split residual 16 crops 32x24"; timeout 200 python tools/trace_halo_problem.py 96 96 9 1 0 1 16 32 24 48 2>&1 | tail -1
echo "=== 256->48 3x3 split 16 crops 64x48"; timeout 200 python tools/trace_halo_problem.py 256 48 9 1 0 0 16 64 48 72 2>&1 | tail -1
echo "=== 64->64 3x3 split 16 crops 64x48"; timeout 200 python tools/trace_halo_problem.py 64 64 9 1 0 0 16 64 48 72 2>&1 | tail -1
echo "=== HRT fc2 320->80 split gelu act-first residual 8 crops"; timeout 200 python tools/trace_halo_problem.py 320 80 1 1 1 1 8 64 48 72 2>&1 | tail -1
for once in 2 1; do
export I2R_HALO_SPLIT_ONCE=$once
for wl in C3 C4 C5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('once=$once $wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
done
