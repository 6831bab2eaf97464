#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s15.log 2>&1
echo "=== new tests"; timeout 900 python -m pytest tests/test_hrnet.py tests/test_compat_extras.py tests/test_hrformer_kernels_gpu.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c45.py -m gpu -q 2>&1 | tail -15
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
echo "=== bench C5"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
echo "=== phase C4 serial"; I2R_CONCURRENT_BRANCHES=0 timeout 300 python tools/phase_times.py coco/interformer_coco_hrt_192_p2_b12.yaml 1 8 > gpurun_out/s15_phase_c4.txt 2>&1; head -1 gpurun_out/s15_phase_c4.txt; tail -14 gpurun_out/s15_phase_c4.txt
