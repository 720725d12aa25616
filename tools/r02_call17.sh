#!/bin/bash
# GPU call 17 (8 GPUs): NCCL check of the sharded winner, bench at N = 8, 4, 2, 1 (C3 incl. the single-stand strong-scaling leg), C4 at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tests/dist_check.py > gpurun_out/r02_c17_distcheck.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/r02_c17_distcheck.log
for n in 8 4 2; do
  timeout 400 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_c17_bench_n$n.json 2> gpurun_out/r02_c17_bench_n$n.err; echo "bench n=$n rc=$?"; tail -2 gpurun_out/r02_c17_bench_n$n.err | cut -c1-300
done
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c17_bench_n1.json 2> gpurun_out/r02_c17_bench_n1.err; echo "bench n=1 rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 > gpurun_out/r02_c17_c4_n8.json 2> gpurun_out/r02_c17_c4_n8.err; echo "c4 n=8 rc=$?"; tail -2 gpurun_out/r02_c17_c4_n8.err | cut -c1-300
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r02_c17_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,2), 'res', round((d.get('e2e_resident_index') or {}).get('value',0)/1e6,2), 'single', {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('single_stand') or {}).items() if k in ('ms','ms_1gpu_same_run','speedup_vs_1gpu','cta_per_icp')}, d['clocks'])
    except Exception as e: print(n, 'FAILED', e)
try:
    d=json.loads(open('gpurun_out/r02_c17_c4_n8.json').read().strip().splitlines()[-1])
    print('c4 n8 value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], 'res', d.get('e2e_resident_index'))
except Exception as e: print('c4 FAILED', e)
PY
