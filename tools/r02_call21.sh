#!/bin/bash
# GPU call 21 (1 GPU): timings of the widened rows (stepwise > 1024 trees, match-and-remove, radial crop), ncu of the grid-build kernels
mkdir -p gpurun_out
timeout 600 python tools/widened_rows_probe.py > gpurun_out/r02_c21_widened.json 2> gpurun_out/r02_c21_widened.err; cat gpurun_out/r02_c21_widened.json; tail -3 gpurun_out/r02_c21_widened.err
CMD="python tools/grid_build_probe.py 2"
$CMD > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"geometry_kernel|bin_kernel|scan_kernel|scatter_kernel|cell_order_kernel" -s 15 -c 5 -f -o gpurun_out/r02_grid_build $CMD > gpurun_out/ncu_grid.log 2>&1
echo "ncu grid rc=$?"
