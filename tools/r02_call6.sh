#!/bin/bash
# GPU call 6: new grid build (one stream-ordered chain, robust extent), skewed-target tests, bench smoke of the new paths
mkdir -p gpurun_out
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_dbg.so timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_next.py -m gpu -x -q > gpurun_out/r02_c6_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c6_dbg.log
tail -15 gpurun_out/r02_c6_dbg.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_c6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c6_pytest.log
tail -25 gpurun_out/r02_c6_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_bench.json').read().strip().splitlines()[-1])
print('value', d['value']/1e6, 'e2e', d['e2e'], 'ms', d['ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','frac_of_hbm_peak')})
print('single', d.get('single_stand'))
print('nn', d.get('nn_query_kernel')); print('grid', d.get('grid_build')); print('clocks', d['clocks'])
PY
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c6_bench_c4.json 2> gpurun_out/r02_c6_bench_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r02_c6_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_bench_c4.json').read().strip().splitlines()[-1])
print('c4 value', d['value']/1e6, 'e2e', d['e2e'], 'ms', d['ms_per_step'], 'grid', d.get('grid_build'), 'launch', d['config']['launch'])
PY
