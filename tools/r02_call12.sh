#!/bin/bash
# GPU call 12: CTA-per-ICP kernel with the shortened trim / fit critical path: debug-assert build, parity, phase clocks, probe
mkdir -p gpurun_out
./tools/microbench/lat > gpurun_out/r02_c12_lat.txt 2>&1; tail -12 gpurun_out/r02_c12_lat.txt
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical" > gpurun_out/r02_c12_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c12_dbg.log
tail -6 gpurun_out/r02_c12_dbg.log
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/r02_c12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c12_pytest.log
tail -6 gpurun_out/r02_c12_pytest.log
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so
for c in 1 0; do timeout 120 python tools/team_phase_clocks.py 8 $c > gpurun_out/r02_c12_clk_w8_c${c}.json 2> gpurun_out/r02_c12_clk.err; cat gpurun_out/r02_c12_clk_w8_c${c}.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c12_clk.err; done
unset FICP_B200_LIB
timeout 300 python tools/strong_scaling_probe.py --worlds 1,2,4,8 --kernels warp,cta,cta1 > gpurun_out/r02_c12_probe.jsonl 2> gpurun_out/r02_c12_probe.err; cut -c1-200 gpurun_out/r02_c12_probe.jsonl; tail -3 gpurun_out/r02_c12_probe.err
timeout 200 python tools/fuzz_parity.py 45 13 > gpurun_out/r02_c12_fuzz.log 2>&1; tail -4 gpurun_out/r02_c12_fuzz.log
