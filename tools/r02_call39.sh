#!/bin/bash
# GPU call 39 (1 GPU): full parity suite at HEAD (call 38 stopped at a Fortran-ordered stack, fixed), config 4 step breakdown
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c39_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c39_pytest.log
tail -3 gpurun_out/r02_c39_pytest.log
timeout 60 python tools/c4_e2e_probe.py > gpurun_out/r02_c4_e2e_probe.jsonl 2> gpurun_out/r02_c4_e2e_probe.err; echo "probe rc=$?"; cat gpurun_out/r02_c4_e2e_probe.jsonl; tail -2 gpurun_out/r02_c4_e2e_probe.err
