#!/bin/bash
# GPU call 22 (1 GPU): warp kernel with the collision pre-check of the fix-up: parity (ICP + trace + soak) and bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/r02_c22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c22_pytest.log
tail -4 gpurun_out/r02_c22_pytest.log
timeout 100 python tools/fuzz_parity.py 40 21 > gpurun_out/r02_c22_fuzz.log 2>&1; tail -2 gpurun_out/r02_c22_fuzz.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-single-stand > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c22_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c22_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'ms', d['ms_per_step'], d['path_stats'])
PY
timeout 300 python bench.py --dims 2 --steps 5 --warmup 3 --no-cpu-baseline --no-single-stand --no-e2e > gpurun_out/r02_c22_bench_xy.json 2> gpurun_out/r02_c22_bench_xy.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c22_bench_xy.json').read().splitlines() if l.startswith('{')][-1])
print('XY value', d['value']/1e6, 'ms', d['ms_per_step'])
PY
