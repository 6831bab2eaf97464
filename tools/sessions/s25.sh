#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s25.log 2>&1
echo "=== s2 tests"; timeout 600 python -m pytest tests/test_halo_s2_gpu.py -m gpu -q 2>&1 | tail -25
echo "=== regression"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['gpu_launches'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
echo "=== phase C2"; timeout 300 python tools/phase_times.py > gpurun_out/s25_phase_c2.txt 2>&1; head -3 gpurun_out/s25_phase_c2.txt; grep -n "k9" gpurun_out/s25_phase_c2.txt | grep -v "48>48k9,32x24:96>48k9" | head -30; tail -12 gpurun_out/s25_phase_c2.txt
