#!/bin/bash
# Final-state validation on one GPU: full parity suite, smoke, bench (default line incl. cpu_baseline), reference arm, launch list,
# ncu --set full of the dominant kernel at HEAD (source of profiles/traffic.json and profiles/r02_icp_kernel_ncu.txt)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r02_final_bench_reference.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_final_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'resident', d['e2e_resident_index']['value']/1e6, 'ms', d['ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','l2_bytes_per_launch','profile_commit_matches_head')}, d['roofline']['on_chip'])
print('single', d.get('single_stand')); print('nn', {k:v for k,v in d['nn_query_kernel'].items() if k in ('ms','queries_per_s','frac_of_l2_peak')}); print('grid', d.get('grid_build')); print('clocks', d['clocks'])
print('cpu', {k:(v if not isinstance(v,dict) else v.get('value',v)) for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
PY
CMD="python bench.py --no-cpu-baseline --no-e2e --no-single-stand --steps 2 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:icp_kernel -s 3 -c 1 -f -o gpurun_out/r02_icp_final $CMD > gpurun_out/ncu_icp.log 2>&1
echo "ncu icp rc=$?"
