#!/bin/bash
# GPU call 2: CTA-per-ICP kernel - debug-assert build first, then the full parity suite, soak, strong-scaling probe
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q -k "cta or trace_is_identical" > gpurun_out/r02_c2_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c2_dbg.log
tail -15 gpurun_out/r02_c2_dbg.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c2_pytest.log
tail -30 gpurun_out/r02_c2_pytest.log
timeout 200 python tools/fuzz_parity.py 90 11 > gpurun_out/r02_c2_fuzz.log 2>&1; tail -8 gpurun_out/r02_c2_fuzz.log
timeout 300 python tools/strong_scaling_probe.py > gpurun_out/r02_c2_probe.jsonl 2> gpurun_out/r02_c2_probe.err; cat gpurun_out/r02_c2_probe.jsonl | cut -c1-400; tail -3 gpurun_out/r02_c2_probe.err
