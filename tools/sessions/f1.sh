#!/bin/bash
# Final single-GPU validation of round 2: full GPU suite, smoke, bench lines of every workload, reference arm, ncu evidence.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/f1.log 2>&1
echo "=== gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --durations=3 2>&1 | tail -6
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== bench C2 (default invocation)"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/r02_bench_c2.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2.json')); print(d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['smem_port'], d.get('cpu_baseline'), d['clocks'])"
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r02_bench_c2_reference_arm.json; cat gpurun_out/r02_bench_c2_reference_arm.json | cut -c1-600
for wl in C3 C4 C5; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl 2>&1 | tail -1 > gpurun_out/r02_bench_$(echo $wl | tr A-Z a-z).json
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_$(echo $wl | tr A-Z a-z).json')); print('$wl', round(d['value'],1), round(d['e2e']['value'],1), round(d['sustained']['value'],1), round(d['roofline']['achieved'],1), d['gpu_launches'])"
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C4 --images 8 2>&1 | tail -1 > gpurun_out/r02_bench_c4_64crops.json
python -c "import json; d=json.load(open('gpurun_out/r02_bench_c4_64crops.json')); print('C4x8', round(d['value'],1), round(d['e2e']['value'],1))"
echo "=== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c2.csv python tools/profile_forward.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_c2.csv | head -24
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_c4.csv python tools/profile_forward.py 1 coco/interformer_coco_hrt_192_p2_b12.yaml 8 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_c4.csv | head -24
echo "=== ncu full: stage-3 BasicBlock groups of conv_halo_kernel"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo_kernel --launch-skip 24 --launch-count 6 -f -o gpurun_out/r02_halo_stage3_full python tools/profile_forward.py > gpurun_out/f1_ncu_full.log 2>&1
tail -3 gpurun_out/f1_ncu_full.log; ls -la gpurun_out/r02_halo_stage3_full.ncu-rep
