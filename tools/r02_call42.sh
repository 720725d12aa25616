#!/bin/bash
# GPU call 42 (1 GPU, the last minute of the round's budget): the tests that exercise the host pass of ficp_batch_create after its
# single-walk rewrite (both row routes, NaN refusal, windows, one-pose batches, goldens)
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_icp.py -x -q -m gpu -k "library_centres or stacked_input or c4_shape or real_data_c1 or window_and_global or golden or tiny_targets" > gpurun_out/r02_c42_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c42_pytest.log
tail -3 gpurun_out/r02_c42_pytest.log
