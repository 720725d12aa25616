# Bench lines of the other workloads / matching modes (no CPU baseline); writes gpurun_out/bench_*.json
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
$B --dims 2 > gpurun_out/bench_xy.json 2>gpurun_out/bench_xy.err
$B --workload c2 > gpurun_out/bench_c2_d3.json 2>gpurun_out/bench_c2_d3.err
$B --workload c2 --dims 2 > gpurun_out/bench_c2_d2.json 2>gpurun_out/bench_c2_d2.err
$B --workload c4 > gpurun_out/bench_c4_d3.json 2>gpurun_out/bench_c4_d3.err
$B --workload c4 --dims 2 > gpurun_out/bench_c4_d2.json 2>gpurun_out/bench_c4_d2.err
