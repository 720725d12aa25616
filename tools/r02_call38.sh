#!/bin/bash
# GPU call 38 (1 GPU, last minutes of the round's budget): parity suite at HEAD after the host-side changes to ficp_batch_create
# (threaded geometry pass, page-locked staging), then config 4 end to end, then the default line without its CPU leg
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c38_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c38_pytest.log
tail -3 gpurun_out/r02_c38_pytest.log
timeout 100 python bench.py --workload c4 --no-cpu-baseline --no-single-stand --steps 5 --warmup 3 --e2e-steps 20 > gpurun_out/r02_c38_bench_c4.json 2> gpurun_out/r02_c38_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c38_bench_c4.json').read().splitlines() if l.startswith('{')][-1])
    print('c4 value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'resident', d['e2e_resident_index'])
except Exception as e:
    print('c4 parse failed', e)
PY
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/r02_c38_bench_c3.json 2> gpurun_out/r02_c38_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c38_bench_c3.json').read().splitlines() if l.startswith('{')][-1])
    print('c3 value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e'].get('ms_per_step_rank0'), 'resident', d['e2e_resident_index']['value']/1e6, 'single', d.get('single_stand',{}).get('ms'))
except Exception as e:
    print('c3 parse failed', e)
PY
