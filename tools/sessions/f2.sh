#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/f2.log 2>&1
echo "=== gpu suite (whole, no -x)"; timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14
