#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s33.log 2>&1
for a in 0 2 4; do
export I2R_HALO_STREAM_ASTG=$a
echo "######## I2R_HALO_STREAM_ASTG=$a"
timeout 300 python tools/chain_probe.py 32 2>&1 | grep "depth 4"
echo "192->96 half"; timeout 200 python tools/trace_halo_problem.py 192 96 9 0 0 1 32 16 12 24 2>&1 | tail -1
echo "256->48"; timeout 200 python tools/trace_halo_problem.py 256 48 9 0 0 0 32 64 48 72 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
