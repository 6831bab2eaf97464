#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s22.log 2>&1
echo "=== chain tests"; timeout 600 python -m pytest tests/test_halo_chain_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -5
for ch in 0 1; do
  I2R_HALO_CHAIN=$ch timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 chain=$ch', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['roofline']['kernel'][:120])"
done
I2R_HALO_CHAIN_COOP=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 coop', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
for ch in 0 1; do
  I2R_HALO_CHAIN=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 chain=$ch', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
echo "=== phase C2 chained"; timeout 300 python tools/phase_times.py > gpurun_out/s22_phase_c2.txt 2>&1; head -40 gpurun_out/s22_phase_c2.txt
