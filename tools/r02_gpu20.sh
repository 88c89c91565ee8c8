#!/bin/bash
# Round 2, GPU call 20: ncu --set full (with source) of the sparse-LBS / GMM / pose / energy kernels of one eager step.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lbs_skin_bwd_kernel|gmm_pass1_kernel|lbs_pose_bwd_warp_kernel|lbs_pose_warp_kernel|fit_losses_kernel|lbs_skin_fwd_kernel' \
  --launch-skip 14 -c 14 -o gpurun_out/r02t_small_kernels -f python tools/run_step_once.py 256 60 2 > gpurun_out/r02t_small_kernels.log 2>&1
ncu -i gpurun_out/r02t_small_kernels.ncu-rep --page details 2>/dev/null | grep -E "^  [a-z_]+.*\(|^    Duration|L2 Cache Throughput|DRAM Throughput|Compute \(SM\) Throughput|Issue Slots Busy|Registers Per Thread|Achieved Occupancy|Theoretical Occupancy|L1/TEX Hit|Warp Cycles Per Issued|Block Limit" | cut -c1-150
