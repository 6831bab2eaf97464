#!/bin/bash
# ncu --set full of the kernels added late in round 2: stride-2 launch of conv_halo_kernel (stem conv2) and window_attention_tc
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/f4.log 2>&1
timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:conv_halo_kernel --launch-count 1 -f -o gpurun_out/r02_halo_s2_stem_conv2 python tools/profile_forward.py > /dev/null 2>&1; ls -la gpurun_out/r02_halo_s2_stem_conv2.ncu-rep
timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:window_attention_tc --launch-count 2 -f -o gpurun_out/r02_window_attention_tc python tools/profile_forward.py 1 coco/interformer_coco_hrt_192_p2_b12.yaml 8 > /dev/null 2>&1; ls -la gpurun_out/r02_window_attention_tc.ncu-rep
