#!/bin/bash
# GPU call 4: CTA-per-ICP kernel v2 (group search, 32-bit sort, parallel fit)
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical" > gpurun_out/r02_c4_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c4_dbg.log
tail -15 gpurun_out/r02_c4_dbg.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_c4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c4_pytest.log
tail -30 gpurun_out/r02_c4_pytest.log
timeout 200 python tools/fuzz_parity.py 60 12 > gpurun_out/r02_c4_fuzz.log 2>&1; tail -8 gpurun_out/r02_c4_fuzz.log
timeout 300 python tools/strong_scaling_probe.py --kernels warp,cta,cta1 > gpurun_out/r02_c4_probe.jsonl 2> gpurun_out/r02_c4_probe.err; cat gpurun_out/r02_c4_probe.jsonl | cut -c1-330; tail -3 gpurun_out/r02_c4_probe.err
CMD="python tools/strong_scaling_probe.py --worlds 8 --kernels cta --reps 1"
$CMD > gpurun_out/plain_team.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_team -s 2 -c 1 -f -o gpurun_out/r02_team_v2 $CMD > gpurun_out/ncu_team.log 2>&1
echo "ncu rc=$?"
