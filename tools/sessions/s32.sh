#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s32.log 2>&1
echo "=== tests"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_hrformer_kernels_gpu.py tests/test_model_gpu_c3.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c45.py -m gpu -q -x 2>&1 | tail -4
for once in 1 0 2; do
export I2R_HALO_SPLIT_ONCE=$once
for wl in C3 C4 C5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('once=$once $wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
done
