#!/bin/bash
# GPU call 16 (1 GPU): C4 launch shapes (warp-per-ICP vs CTA-per-ICP), bench with the resident-index e2e leg
mkdir -p gpurun_out
for mode in auto on; do
  timeout 300 python bench.py --workload c4 --cta $mode --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_c4_$mode.json 2> gpurun_out/r02_c16_c4.err; echo "c4 $mode rc=$?"; tail -2 gpurun_out/r02_c16_c4.err
  python - $mode <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/r02_c16_c4_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('c4', sys.argv[1], 'value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'resident', d['e2e_resident_index'], 'launch', {k:d['config']['launch'][k] for k in ('cta_per_icp','warps_per_cta','ctas','ctas_per_sm','smem_bytes','elems_per_lane')})
PY
done
timeout 300 python bench.py --workload c4 --dims 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_c4_xy.json 2> gpurun_out/r02_c16_c4.err; echo "c4 xy rc=$?"
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_c5.json 2> gpurun_out/r02_c16_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/r02_c16_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c16_c5.json').read().strip().splitlines()[-1])
print('c5 value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']); print(d.get('trim_fraction_sweep'))
PY
