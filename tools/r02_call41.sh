#!/bin/bash
# GPU call 41 (1 GPU, the round's last GPU minutes): parity suite at HEAD (page-locked rows uploaded as they are and split on the
# device), config 4 step breakdown by memory kind, config 4 line
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/r02_c41_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c41_pytest.log
tail -3 gpurun_out/r02_c41_pytest.log
timeout 40 python tools/c4_e2e_probe.py > gpurun_out/r02_c4_e2e_probe_c41.jsonl 2> gpurun_out/r02_c4_e2e_probe_c41.err; echo "probe rc=$?"; cut -c1-400 gpurun_out/r02_c4_e2e_probe_c41.jsonl; tail -2 gpurun_out/r02_c4_e2e_probe_c41.err
timeout 60 python bench.py --workload c4 --no-cpu-baseline --no-single-stand --steps 5 --warmup 3 --e2e-steps 20 > gpurun_out/r02_c41_bench_c4.json 2> gpurun_out/r02_c41_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c41_bench_c4.json').read().splitlines() if l.startswith('{')][-1])
    r=d['e2e_resident_index']
    print('c4 value', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'resident', round(r['value']/1e6,3), 'stacked', round(r.get('stacked_input',{}).get('value',0)/1e6,3))
except Exception as e:
    print('parse failed', e)
PY
