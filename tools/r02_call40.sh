#!/bin/bash
# GPU call 40 (1 GPU): parity suite at HEAD (library-side centres, sliced upload), config 4 step breakdown, config 4 and default lines
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r02_c40_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c40_pytest.log
tail -3 gpurun_out/r02_c40_pytest.log
timeout 60 python tools/c4_e2e_probe.py > gpurun_out/r02_c4_e2e_probe_head.jsonl 2> gpurun_out/r02_c4_e2e_probe_head.err; echo "probe rc=$?"; cut -c1-420 gpurun_out/r02_c4_e2e_probe_head.jsonl
timeout 100 python bench.py --workload c4 --no-cpu-baseline --no-single-stand --steps 5 --warmup 3 --e2e-steps 20 > gpurun_out/r02_c40_bench_c4.json 2> gpurun_out/r02_c40_bench_c4.err; echo "c4 rc=$?"
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/r02_c40_bench_c3.json 2> gpurun_out/r02_c40_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for w in ("c4", "c3"):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r02_c40_bench_{w}.json').read().splitlines() if l.startswith('{')][-1])
        r=d['e2e_resident_index']
        print(w, 'value', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'resident', round(r['value']/1e6,3), 'stacked', round(r.get('stacked_input',{}).get('value',0)/1e6,3), 'single', d.get('single_stand',{}).get('ms'))
    except Exception as e:
        print(w, 'parse failed', e)
PY
