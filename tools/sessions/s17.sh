#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s17.log 2>&1
echo "=== sanity e12"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_hrformer_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "=== sanity e16"; I2R_LIB=build/libi2r_e16.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for v in "" build/libi2r_e16.so build/libi2r_e8.so; do
  echo "=== variant [$v]"
  for wl in C2 C3 C4; do
    I2R_LIB=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1), round(d['roofline']['achieved'],1))"
  done
done
echo "=== fc1 trace e12"; timeout 200 python tools/trace_halo_problem.py 80 320 1 1 1 0 8 64 48 12 2>&1 | tail -16
