#!/bin/bash
# GPU call 10: bulk NN kernel v2 (records + plan kernel + ring-finish kernel), latency micro-benchmarks, A/B of the own-cell prescan
mkdir -p gpurun_out
./tools/microbench/lat > gpurun_out/r02_c10_lat.txt 2>&1; cat gpurun_out/r02_c10_lat.txt
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nn_query" > gpurun_out/r02_c10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c10_pytest.log
tail -15 gpurun_out/r02_c10_pytest.log
timeout 300 python tools/nn_bulk_probe.py 3 > gpurun_out/r02_c10_bulk3.jsonl 2> gpurun_out/r02_c10_bulk3.err; cut -c1-330 gpurun_out/r02_c10_bulk3.jsonl; tail -3 gpurun_out/r02_c10_bulk3.err
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_b3.so timeout 300 python tools/nn_bulk_probe.py 3 > gpurun_out/r02_c10_bulk3_b3.jsonl 2> gpurun_out/r02_c10_bulk3_b3.err; cut -c1-330 gpurun_out/r02_c10_bulk3_b3.jsonl; tail -3 gpurun_out/r02_c10_bulk3_b3.err
timeout 300 python tools/nn_bulk_probe.py 2 > gpurun_out/r02_c10_bulk2.jsonl 2> gpurun_out/r02_c10_bulk2.err; cut -c1-330 gpurun_out/r02_c10_bulk2.jsonl; tail -3 gpurun_out/r02_c10_bulk2.err
bash tools/ab_variants.sh b200 pre 2>&1 | tee gpurun_out/r02_c10_ab.log
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_pre.so timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/r02_c10_pytest_pre.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c10_pytest_pre.log
tail -5 gpurun_out/r02_c10_pytest_pre.log
CMD="python tools/nn_bulk_probe.py 3 22"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_c10_bulk_launches.csv $CMD > gpurun_out/ncu_bl.log 2>&1
echo "launch list rc=$?"
# the 4th density (3 points per cell): launches of nn_bulk_kernel come 8 per density -> skip 3*8+2
ncu --set full --clock-control none --import-source on -k regex:nn_bulk_kernel -s 26 -c 1 -f -o gpurun_out/r02_nn_bulk_v2 $CMD > gpurun_out/ncu_bulk.log 2>&1
echo "ncu bulk rc=$?"
