#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s13.log 2>&1
echo "=== stress"; timeout 600 python -m pytest tests/test_halo_stress_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "=== trace pair on"; STEP=1 timeout 300 python tools/trace_halo_summary.py --res 2>&1 | tail -150 > gpurun_out/s13_trace_pair.txt; awk 'NR%4==1' gpurun_out/s13_trace_pair.txt | head -45
echo "=== trace pair off"; I2R_HALO_PAIR=0 STEP=3 timeout 300 python tools/trace_halo_summary.py --res 2>&1 | tail -52 > gpurun_out/s13_trace_nopair.txt; awk 'NR%3==1' gpurun_out/s13_trace_nopair.txt | head -20
echo "=== phase times pair on"; timeout 300 python tools/phase_times.py 2>&1 | tail -102 > gpurun_out/s13_phase_pair.txt; head -1 gpurun_out/s13_phase_pair.txt
echo "=== phase times pair off"; I2R_HALO_PAIR=0 timeout 300 python tools/phase_times.py 2>&1 | tail -102 > gpurun_out/s13_phase_nopair.txt; head -1 gpurun_out/s13_phase_nopair.txt
echo "=== bench C2 pair on"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel'][:160])"
echo "=== bench C2 pair off"; I2R_HALO_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel'][:160])"
echo "=== bench C3 pair on"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"
echo "=== bench C3 pair off"; I2R_HALO_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.5 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"
