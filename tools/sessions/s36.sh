#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s36.log 2>&1
echo "=== compute-sanitizer memcheck: new kernels of round 2 (stride-2 halo, chained launches, window attention, post/pre-processing, mask res stem)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 --launch-timeout 0 python -m pytest tests/test_halo_s2_gpu.py tests/test_halo_chain_gpu.py tests/test_postproc.py tests/test_preproc.py -m gpu -q -x -k "not c2_forward and not vanilla_forward and not forward_flip" 2>&1 | tail -15
echo "rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 --launch-timeout 0 python -m pytest tests/test_hrformer_kernels_gpu.py -m gpu -q -x -k "window_attention" 2>&1 | tail -8
