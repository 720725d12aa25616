#!/bin/bash
# GPU call 14: CTA-per-ICP kernel with CTA-uniform state in shared memory (fewer spills at 64 registers)
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical" > gpurun_out/r02_c14_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c14_dbg.log
tail -6 gpurun_out/r02_c14_dbg.log
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so
for c in 1 0; do timeout 120 python tools/team_phase_clocks.py 8 $c > gpurun_out/r02_c14_clk_w8_c${c}.json 2> gpurun_out/r02_c14_clk.err; cat gpurun_out/r02_c14_clk_w8_c${c}.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c14_clk.err; done
unset FICP_B200_LIB
timeout 300 python tools/strong_scaling_probe.py --worlds 1,8 --kernels cta,cta1 > gpurun_out/r02_c14_probe.jsonl 2> gpurun_out/r02_c14_probe.err; cut -c1-420 gpurun_out/r02_c14_probe.jsonl; tail -3 gpurun_out/r02_c14_probe.err
