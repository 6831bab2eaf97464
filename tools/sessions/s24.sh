#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s24.log 2>&1
echo "=== chain tests"; timeout 600 python -m pytest tests/test_halo_chain_gpu.py tests/test_halo_stress_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 > gpurun_out/s24_c2.json
python -c "
import json; d=json.load(open('gpurun_out/s24_c2.json')); print('C2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['roofline']['smem_port'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['roofline']['smem_port'])"
