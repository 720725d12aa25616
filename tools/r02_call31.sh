#!/bin/bash
# GPU call 31 (2 GPUs): shard_plan on the shipped path - NCCL winner check (hypothesis cut, plot cut with one pose, plot cut with many poses), bench at N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/r02_c31_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/r02_c31_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c31_bench_n2.json 2> gpurun_out/r02_c31_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_c31_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c31_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'resident', d['e2e_resident_index']['value']/1e6, d['config']['parallelism'])
print('single', d.get('single_stand'))
PY
