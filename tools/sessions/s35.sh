#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s35.log 2>&1
echo "=== tests"; timeout 900 python -m pytest tests/test_hrformer_kernels_gpu.py tests/test_model_gpu_hrt.py tests/test_model_gpu_c45.py -m gpu -q -x 2>&1 | tail -4
for k in tc mma; do
export I2R_WINDOW_ATT=$k
for wl in C4 C5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('att=$k $wl', round(d['value'],1), round(d['e2e']['value'],1))"
done
done
unset I2R_WINDOW_ATT
echo "=== phase C4 serial"; I2R_CONCURRENT_BRANCHES=0 timeout 300 python tools/phase_times.py coco/interformer_coco_hrt_192_p2_b12.yaml 1 8 > gpurun_out/s35_phase_c4.txt 2>&1; head -1 gpurun_out/s35_phase_c4.txt; sed -n 12,22p gpurun_out/s35_phase_c4.txt; tail -14 gpurun_out/s35_phase_c4.txt
