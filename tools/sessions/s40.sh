#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s40.log 2>&1
for rep in 1 2; do
for v in A B; do
if [ $v = A ]; then export I2R_LIB=$GRAFT_REPO_ROOT/build/libi2r_final_a.so; else unset I2R_LIB; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v C2', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1))"
done
done
unset I2R_LIB
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustain-seconds 0.3 --workload C4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B C4', round(d['value'],1))"
echo "=== gpu suite on B"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
