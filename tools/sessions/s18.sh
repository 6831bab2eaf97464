#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s19.log 2>&1
echo "=== kernels + stress"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_halo_stress_gpu.py tests/test_hrformer_kernels_gpu.py tests/test_attention_tc_gpu.py -m gpu -q -x 2>&1 | tail -12
echo "=== models"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_gpu_c3.py tests/test_model_gpu_hrt.py tests/test_halo_pair_gpu.py -m gpu -q 2>&1 | tail -8
for st in 1 0; do
  echo "=== stage=$st"
  for wl in C2 C3 C4; do
    I2R_HALO_STAGE=$st timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1), round(d['roofline']['achieved'],1))"
  done
done
echo "=== trace stage3 group"; STEP=6 timeout 300 python tools/trace_halo_summary.py --res 2>&1 | tail -27
