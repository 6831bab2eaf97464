#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s28.log 2>&1
echo "=== tests"; timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "=== layer1 conv3: 64->256 1x1 + residual, 32 crops"; timeout 200 python tools/trace_halo_problem.py 64 256 1 0 0 1 32 64 48 48 2>&1 | tail -5
echo "=== same, res_tma off"; I2R_HALO_RES_TMA=0 timeout 200 python tools/trace_halo_problem.py 64 256 1 0 0 1 32 64 48 48 2>&1 | tail -5
echo "=== HRT fc2 320->80 split gelu act-first residual 8 crops"; timeout 200 python tools/trace_halo_problem.py 320 80 1 1 1 1 8 64 48 48 2>&1 | tail -5
echo "=== same, res_tma everywhere"; I2R_HALO_RES_TMA=2 timeout 200 python tools/trace_halo_problem.py 320 80 1 1 1 1 8 64 48 48 2>&1 | tail -5
for wl in C2 C4; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
I2R_HALO_RES_TMA=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 res_tma=2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
