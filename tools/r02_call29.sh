#!/bin/bash
# GPU call 29 (1 GPU): CTA-per-ICP kernel, one vs two CTAs per SM on the 8- and 4-rank shards of the single stand; phase clocks alone on the SM
mkdir -p gpurun_out
timeout 300 python tools/strong_scaling_probe.py --worlds 4,8 --kernels cta,cta1 --reps 7 > gpurun_out/r02_c29_probe.jsonl 2> gpurun_out/r02_c29_probe.err; cut -c1-200 gpurun_out/r02_c29_probe.jsonl; tail -2 gpurun_out/r02_c29_probe.err
for c in 1 2; do
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so timeout 120 python tools/team_phase_clocks.py 8 $c > gpurun_out/r02_c29_clk_$c.json 2> gpurun_out/r02_c29_clk.err; cat gpurun_out/r02_c29_clk_$c.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c29_clk.err
done
