#!/bin/bash
# GPU call 3: ncu of the CTA-per-ICP kernel on the 8-GPU shard of the literal C3 stand
mkdir -p gpurun_out
CMD="python tools/strong_scaling_probe.py --worlds 8 --kernels cta --reps 1"
$CMD > gpurun_out/plain_team.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_team -s 2 -c 1 -f -o gpurun_out/r02_team_v1 $CMD > gpurun_out/ncu_team.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_team.log | cut -c1-300
