#!/bin/bash
# GPU call 15: full parity suite at HEAD, bench line, repair-threshold A/B of the CTA-per-ICP kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c15_pytest.log
tail -6 gpurun_out/r02_c15_pytest.log
for v in b200 rep48 rep128; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so timeout 300 python tools/strong_scaling_probe.py --worlds 8 --kernels cta,cta1 --reps 9 > gpurun_out/r02_c15_probe_$v.jsonl 2> gpurun_out/r02_c15_probe.err; cut -c1-330 gpurun_out/r02_c15_probe_$v.jsonl; tail -2 gpurun_out/r02_c15_probe.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_c15_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c15_bench.json').read().strip().splitlines()[-1])
print('value', d['value']/1e6, 'e2e', d['e2e'], 'ms', d['ms_per_step'])
print('single', d.get('single_stand'))
print('nn', d.get('nn_query_kernel')); print('grid', d.get('grid_build'))
PY
