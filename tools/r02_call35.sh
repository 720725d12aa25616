#!/bin/bash
# GPU call 35 (1 GPU): device-resident stepper for plots above 1024 trees - parity with the host-stepped path, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_stages.py tests/test_reference_vendored.py -m gpu -x -q > gpurun_out/r02_c35_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c35_pytest.log; tail -4 gpurun_out/r02_c35_pytest.log
timeout 600 python tools/large_plot_probe.py > gpurun_out/r02_c35_large_plot.jsonl 2> gpurun_out/r02_c35_large_plot.err; cat gpurun_out/r02_c35_large_plot.jsonl; tail -2 gpurun_out/r02_c35_large_plot.err
