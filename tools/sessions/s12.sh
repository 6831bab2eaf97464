#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s12.log 2>&1
echo "=== phase times pair on"; timeout 300 python tools/phase_times.py 2>&1 | tail -102 > gpurun_out/s12_phase_pair.txt; head -1 gpurun_out/s12_phase_pair.txt
echo "=== phase times pair off"; I2R_HALO_PAIR=0 timeout 300 python tools/phase_times.py 2>&1 | tail -102 > gpurun_out/s12_phase_nopair.txt; head -1 gpurun_out/s12_phase_nopair.txt
echo "=== trace pair on"; STEP=1 timeout 300 python tools/trace_halo_summary.py --res 2>&1 | tail -150
echo "=== trace pair off"; I2R_HALO_PAIR=0 STEP=3 timeout 300 python tools/trace_halo_summary.py --res 2>&1 | tail -55
