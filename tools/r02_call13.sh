#!/bin/bash
# GPU call 13: CTA-per-ICP kernel compiled for 128 registers (no spills) vs 64: phase clocks and probe
mkdir -p gpurun_out
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_r128.so
timeout 120 python tools/team_phase_clocks.py 8 0 > gpurun_out/r02_c13_clk_r128.json 2> gpurun_out/r02_c13_clk.err; cat gpurun_out/r02_c13_clk_r128.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c13_clk.err
timeout 300 python tools/strong_scaling_probe.py --worlds 1,4,8 --kernels cta > gpurun_out/r02_c13_probe_r128.jsonl 2> gpurun_out/r02_c13_probe.err; cut -c1-400 gpurun_out/r02_c13_probe_r128.jsonl; tail -3 gpurun_out/r02_c13_probe.err
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so
timeout 120 python tools/team_phase_clocks.py 8 1 > gpurun_out/r02_c13_clk_r64.json 2> gpurun_out/r02_c13_clk.err; cat gpurun_out/r02_c13_clk_r64.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c13_clk.err
