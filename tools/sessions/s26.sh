#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s26.log 2>&1
echo "=== layer1 conv3: 64->256 1x1 + residual, 32 crops"; timeout 200 python tools/trace_halo_problem.py 64 256 1 0 0 1 32 64 48 24 2>&1 | tail -14
echo "=== layer1 conv1: 256->64 1x1, 32 crops"; timeout 200 python tools/trace_halo_problem.py 256 64 1 0 0 0 32 64 48 24 2>&1 | tail -12
echo "=== transition 256->48 3x3"; timeout 200 python tools/trace_halo_problem.py 256 48 9 0 0 0 32 64 48 24 2>&1 | tail -12
