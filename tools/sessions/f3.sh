#!/bin/bash
# Final single-GPU validation (lean: no ncu --set full capture; outputs stay small)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/f3.log 2>&1
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench C2 (default invocation)"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/r02_bench_c2.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['smem_port']['frac'], d.get('cpu_baseline',{}).get('value'), d['clocks'], d['gpu_launches'])"
for wl in C3 C4 C5; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl 2>&1 | tail -1 > gpurun_out/r02_bench_$(echo $wl | tr A-Z a-z).json
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_$(echo $wl | tr A-Z a-z).json')); print('$wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['sustained']['value'],1), round(d['roofline']['achieved'],1), d['gpu_launches'])"
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 --images 8 2>&1 | tail -1 > gpurun_out/r02_bench_c4_64crops.json
python -c "import json; d=json.load(open('gpurun_out/r02_bench_c4_64crops.json')); print('C4x8', round(d['value'],1), round(d['e2e']['value'],1))"
echo "=== ncu launch list C2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c2.csv python tools/profile_forward.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_c2.csv | head -14
