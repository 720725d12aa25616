#!/bin/bash
# GPU call 11: bulk NN kernel v3 (run loops, exact per-row plans), ncu source-level capture of the CTA-per-ICP kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nn_query" > gpurun_out/r02_c11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c11_pytest.log
tail -8 gpurun_out/r02_c11_pytest.log
timeout 300 python tools/nn_bulk_probe.py 3 > gpurun_out/r02_c11_bulk3.jsonl 2> gpurun_out/r02_c11_bulk3.err; cut -c1-330 gpurun_out/r02_c11_bulk3.jsonl; tail -3 gpurun_out/r02_c11_bulk3.err
timeout 300 python tools/nn_bulk_probe.py 2 > gpurun_out/r02_c11_bulk2.jsonl 2> gpurun_out/r02_c11_bulk2.err; cut -c1-330 gpurun_out/r02_c11_bulk2.jsonl; tail -3 gpurun_out/r02_c11_bulk2.err
CMD="python tools/strong_scaling_probe.py --worlds 8 --kernels cta1 --reps 1"
$CMD > gpurun_out/plain_team.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_team -s 2 -c 1 -f -o gpurun_out/r02_team_head $CMD > gpurun_out/ncu_team.log 2>&1
echo "ncu team rc=$?"
