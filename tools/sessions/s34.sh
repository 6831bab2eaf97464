#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s34.log 2>&1
echo "=== window attention default desc (lbo 16, sbo 1024, MN-major)"; timeout 300 python -m pytest tests/test_hrformer_kernels_gpu.py -m gpu -q -k window_attention 2>&1 | tail -12
for d in "1024,16,1" "16,1024,0" "0,1024,1" "2048,1024,1"; do
echo "=== I2R_WATT_DESC=$d"; I2R_WATT_DESC=$d timeout 300 python -m pytest tests/test_hrformer_kernels_gpu.py -m gpu -q -k "window_attention and tcgen05" 2>&1 | tail -3
done
