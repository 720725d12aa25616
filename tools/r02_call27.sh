#!/bin/bash
# GPU call 27 (1 GPU): CTA-per-ICP kernel with two trees per thread (128 registers, 2 CTAs/SM): debug-assert build, parity, probe, clocks
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical or planner" > gpurun_out/r02_c27_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c27_dbg.log
tail -6 gpurun_out/r02_c27_dbg.log
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/r02_c27_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c27_pytest.log
tail -6 gpurun_out/r02_c27_pytest.log
timeout 100 python tools/fuzz_parity.py 45 31 > gpurun_out/r02_c27_fuzz.log 2>&1; tail -2 gpurun_out/r02_c27_fuzz.log
for v in b200 tpt1; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so timeout 300 python tools/strong_scaling_probe.py --worlds 1,2,4,8 --kernels cta,warp --reps 7 > gpurun_out/r02_c27_probe_$v.jsonl 2> gpurun_out/r02_c27_probe.err; cut -c1-230 gpurun_out/r02_c27_probe_$v.jsonl; tail -2 gpurun_out/r02_c27_probe.err
done
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so timeout 120 python tools/team_phase_clocks.py 8 0 > gpurun_out/r02_c27_clk.json 2> gpurun_out/r02_c27_clk.err; cat gpurun_out/r02_c27_clk.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c27_clk.err
