#!/bin/bash
# GPU call 8 (re-entry): parity of HEAD, bench, launch list, ncu --set full of the HEAD icp_kernel / nn_query kernel, probes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c8_pytest.log
tail -8 gpurun_out/r02_c8_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c8_bench.json').read().strip().splitlines()[-1])
print('value', d['value']/1e6, 'e2e', d['e2e'], 'ms', d['ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','frac_of_hbm_peak')})
print('single', d.get('single_stand'))
print('nn', d.get('nn_query_kernel')); print('grid', d.get('grid_build')); print('clocks', d['clocks'])
print('cpu', d.get('cpu_baseline'))
PY
timeout 300 python tools/grid_build_probe.py 5 > gpurun_out/r02_c8_grid.jsonl 2> gpurun_out/r02_c8_grid.err; cat gpurun_out/r02_c8_grid.jsonl; tail -3 gpurun_out/r02_c8_grid.err
timeout 300 python tools/strong_scaling_probe.py --worlds 1,2,4,8 --kernels warp,cta,auto > gpurun_out/r02_c8_probe.jsonl 2> gpurun_out/r02_c8_probe.err; cut -c1-200 gpurun_out/r02_c8_probe.jsonl; tail -3 gpurun_out/r02_c8_probe.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c8_bench_c4.json 2> gpurun_out/r02_c8_bench_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r02_c8_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c8_bench_c4.json').read().strip().splitlines()[-1])
print('c4 value', d['value']/1e6, 'e2e', d['e2e'], 'ms', d['ms_per_step'], 'grid', d.get('grid_build'), 'launch', d['config']['launch'])
PY
CMD="python bench.py --no-cpu-baseline --no-e2e --no-single-stand --steps 2 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c8_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:icp_kernel -s 3 -c 1 -f -o gpurun_out/r02_icp_head $CMD > gpurun_out/ncu_icp.log 2>&1
echo "ncu icp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:nn_query -s 1 -c 1 -f -o gpurun_out/r02_nnq_head $CMD > gpurun_out/ncu_nnq.log 2>&1
echo "ncu nnq rc=$?"
ls -la gpurun_out/*.ncu-rep
