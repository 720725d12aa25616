#!/bin/bash
# GPU call 5: CTA-per-ICP kernel v3 (previous trim order first, warp-0 scan)
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical" > gpurun_out/r02_c5_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c5_dbg.log
tail -15 gpurun_out/r02_c5_dbg.log
timeout 300 python tools/strong_scaling_probe.py --kernels cta,cta1 --worlds 1,4,8 > gpurun_out/r02_c5_probe.jsonl 2> gpurun_out/r02_c5_probe.err; cat gpurun_out/r02_c5_probe.jsonl | cut -c1-330; tail -3 gpurun_out/r02_c5_probe.err
timeout 200 python tools/fuzz_parity.py 45 13 > gpurun_out/r02_c5_fuzz.log 2>&1; tail -4 gpurun_out/r02_c5_fuzz.log
CMD="python tools/strong_scaling_probe.py --worlds 8 --kernels cta --reps 1"
$CMD > gpurun_out/plain_team.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_team -s 2 -c 1 -f -o gpurun_out/r02_team_v3 $CMD > gpurun_out/ncu_team.log 2>&1
echo "ncu rc=$?"
