#!/bin/bash
# GPU call 28 (1 GPU): HEAD validation - full parity suite (incl. the large-plot trim path), soak regression (seed 31), smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c28_pytest.log
tail -8 gpurun_out/r02_c28_pytest.log
timeout 100 python tools/fuzz_parity.py 45 31 > gpurun_out/r02_c28_fuzz.log 2>&1; tail -2 gpurun_out/r02_c28_fuzz.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c28_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_c28_smoke.log
timeout 900 python bench.py > gpurun_out/r02_c28_bench.json 2> gpurun_out/r02_c28_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c28_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c28_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'resident', d['e2e_resident_index']['value']/1e6, 'ms', d['ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','l2_bytes_per_launch','profile_commit_matches_head')}, d['roofline'].get('on_chip'))
print('single', d.get('single_stand')); print('clocks', d['clocks'])
print('cpu', {k:(v if not isinstance(v,dict) else v.get('value',v)) for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
PY
