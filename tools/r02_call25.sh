#!/bin/bash
# GPU call 25 (1 GPU): longer randomised soak at HEAD (two seeds), C2 bench lines, C3 XY
mkdir -p gpurun_out
for seed in 101 202; do timeout 200 python tools/fuzz_parity.py 120 $seed > gpurun_out/r02_c25_fuzz_$seed.log 2>&1; tail -1 gpurun_out/r02_c25_fuzz_$seed.log; done
for d in 3 2; do
  timeout 300 python bench.py --workload c2 --dims $d --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c25_c2_d$d.json 2> gpurun_out/r02_c25_c2.err; tail -1 gpurun_out/r02_c25_c2.err
done
timeout 300 python bench.py --dims 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c25_c3_xy.json 2> gpurun_out/r02_c25_c3.err
python - <<'PY'
import json
for f in ('r02_c25_c2_d3','r02_c25_c2_d2','r02_c25_c3_xy'):
    d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().splitlines() if l.startswith('{')][-1])
    print(f, 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'ms', round(d['ms_per_step'],2), 'single', (d.get('single_stand') or {}).get('ms'), d['config']['launch']['cta_per_icp'])
PY
