#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s5.log 2>&1
echo "=== stress -x"; timeout 900 python -m pytest tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -60
echo "=== hrt 64 crops ungrouped"; timeout 300 python tools/hang_hunt_model.py coco/interformer_coco_hrt_192_p2_b12.yaml 8 8 2>&1 | tail -25
echo "=== hrt 24 crops"; timeout 300 python tools/hang_hunt_model.py coco/interformer_coco_hrt_192_p2_b12.yaml 8 3 2>&1 | tail -25
echo "=== hrt ragged golden"; timeout 300 python -m pytest "tests/test_model_gpu_hrt.py" -m gpu -q -x 2>&1 | tail -30
