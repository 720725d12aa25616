# A/B helper: launch-shape variants of the default workload (no CPU baseline)
B="python bench.py --no-cpu-baseline --no-e2e --no-single-stand --steps 3 --warmup 3"
$B --helpers on > gpurun_out/ab_helpers.json 2>gpurun_out/ab_helpers.err
$B --plots-per-gpu 32 > gpurun_out/ab_p32.json 2>gpurun_out/ab_p32.err
$B --plots-per-gpu 32 --helpers on > gpurun_out/ab_p32_helpers.json 2>gpurun_out/ab_p32_helpers.err
