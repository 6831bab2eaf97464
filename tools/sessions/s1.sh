#!/bin/bash
# GPU session 1: hang hunt (original faulting build vs instrumented build) + full GPU suite + new C4/C5 parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s1.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))"
for cfg in "24 64 256 1 1 1" "48 64 256 1 1 1" "24 96 96 1 1 1 9" "64 192 192 0 1 1 9"; do
  echo "=== orig $cfg"; I2R_LIB=build/libi2r_bad_orig.so timeout 180 python tools/hang_hunt.py $cfg 2>&1 | tail -5
  echo "=== new  $cfg"; timeout 180 python tools/hang_hunt.py $cfg 2>&1 | tail -25
done
echo "=== new gelu"; timeout 180 python tools/hang_hunt.py 24 64 256 1 1 0 1 3 1 2>&1 | tail -25
echo "=== pytest (old suite)"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_model_gpu_c45.py 2>&1 | tail -15
echo "=== pytest c45"
timeout 1200 python -m pytest tests/test_model_gpu_c45.py -m gpu -q 2>&1 | tail -30
cat gpurun_out/model_report.jsonl
