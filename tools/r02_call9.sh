#!/bin/bash
# GPU call 9: bulk NN kernel (parity + probe over grid densities), phase clocks of the CTA-per-ICP kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nn_query" > gpurun_out/r02_c9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c9_pytest.log
tail -15 gpurun_out/r02_c9_pytest.log
timeout 300 python tools/nn_bulk_probe.py 3 > gpurun_out/r02_c9_bulk3.jsonl 2> gpurun_out/r02_c9_bulk3.err; cat gpurun_out/r02_c9_bulk3.jsonl; tail -3 gpurun_out/r02_c9_bulk3.err
timeout 300 python tools/nn_bulk_probe.py 2 > gpurun_out/r02_c9_bulk2.jsonl 2> gpurun_out/r02_c9_bulk2.err; cat gpurun_out/r02_c9_bulk2.jsonl; tail -3 gpurun_out/r02_c9_bulk2.err
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so
for w in 8 1; do for c in 0 1; do timeout 120 python tools/team_phase_clocks.py $w $c > gpurun_out/r02_c9_clk_w${w}_c${c}.json 2> gpurun_out/r02_c9_clk.err; cat gpurun_out/r02_c9_clk_w${w}_c${c}.json | tr -d '\n ' ; echo; tail -2 gpurun_out/r02_c9_clk.err; done; done
unset FICP_B200_LIB
CMD="python tools/nn_bulk_probe.py 3 22"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c9_bulk_launches.csv $CMD > gpurun_out/ncu_bl.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:nn_bulk_kernel -s 6 -c 1 -f -o gpurun_out/r02_nn_bulk $CMD > gpurun_out/ncu_bulk.log 2>&1
echo "ncu bulk rc=$?"
