#!/bin/bash
mkdir -p gpurun_out
python tools/grid_build_probe.py 5 > gpurun_out/r02_c7_grid.jsonl 2> gpurun_out/r02_c7_grid.err; cat gpurun_out/r02_c7_grid.jsonl; tail -3 gpurun_out/r02_c7_grid.err
python tools/grid_build_probe.py 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c7_grid_launches.csv python tools/grid_build_probe.py 1 > gpurun_out/ncu_grid.log 2>&1
echo "ncu rc=$?"
