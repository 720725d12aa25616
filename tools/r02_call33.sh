#!/bin/bash
# GPU call 33 (1 GPU): ncu --set full of the shipped CTA-per-ICP kernel (two trees per thread) on the 8-rank shard of the single stand
mkdir -p gpurun_out
CMD="python tools/strong_scaling_probe.py --worlds 8 --kernels cta --reps 3"
$CMD > gpurun_out/r02_c33_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_team_kernel -s 2 -c 1 -f -o gpurun_out/r02_team_head $CMD > gpurun_out/r02_c33_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_c33_ncu.log; ls -la gpurun_out/r02_team_head.ncu-rep
