#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s30.log 2>&1
for once in 1 0; do
export I2R_HALO_SPLIT_ONCE=$once
echo "######## I2R_HALO_SPLIT_ONCE=$once"
echo "=== 192->192 3x3 split residual 16 crops 16x12"; timeout 200 python tools/trace_halo_problem.py 192 192 9 1 0 1 16 16 12 16 2>&1 | tail -2
echo "=== 96->96 3x3 split residual 16 crops 32x24"; timeout 200 python tools/trace_halo_problem.py 96 96 9 1 0 1 16 32 24 48 2>&1 | tail -2
echo "=== 48->48 3x3 split residual 16 crops 64x48"; timeout 200 python tools/trace_halo_problem.py 48 48 9 1 0 1 16 64 48 72 2>&1 | tail -2
echo "=== 256->48 3x3 split 16 crops 64x48"; timeout 200 python tools/trace_halo_problem.py 256 48 9 1 0 0 16 64 48 72 2>&1 | tail -2
echo "=== HRT fc2 320->80 split gelu act-first residual 8 crops"; timeout 200 python tools/trace_halo_problem.py 320 80 1 1 1 1 8 64 48 72 2>&1 | tail -2
echo "=== HRT fc1 80->320 split gelu 8 crops"; timeout 200 python tools/trace_halo_problem.py 80 320 1 1 1 0 8 64 48 72 2>&1 | tail -2
done
