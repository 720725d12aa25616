#!/bin/bash
# GPU call 18 (1 GPU): new tests (packed world translation, content-keyed index cache), C4 e2e after the host-prep vectorisation
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_stages.py -m gpu -x -q -k "packed_world or edited_in_place or batch" > gpurun_out/r02_c18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c18_pytest.log
tail -5 gpurun_out/r02_c18_pytest.log
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c18_c4.json 2> gpurun_out/r02_c18_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r02_c18_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c18_c4.json').read().strip().splitlines()[-1])
print('c4 value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'resident', d['e2e_resident_index'], 'launch', {k:d['config']['launch'][k] for k in ('cta_per_icp','warps_per_cta','ctas','ctas_per_sm')})
PY
timeout 120 python tools/e2e_breakdown.py
