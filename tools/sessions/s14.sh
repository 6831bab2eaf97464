#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s14.log 2>&1
echo "=== preproc + pair + hrt tests"; timeout 900 python -m pytest tests/test_preproc.py tests/test_halo_pair_gpu.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c45.py -m gpu -q 2>&1 | tail -12
echo "=== phase C4 serial"; I2R_CONCURRENT_BRANCHES=0 timeout 300 python tools/phase_times.py coco/interformer_coco_hrt_192_p2_b12.yaml 1 8 > gpurun_out/s14_phase_c4.txt 2>&1; head -1 gpurun_out/s14_phase_c4.txt; tail -14 gpurun_out/s14_phase_c4.txt
echo "=== bench C4"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
echo "=== bench C5"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
echo "=== bench C4 x8 images"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C4 --images 8 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
