#!/bin/bash
# GPU call 32 (8 GPUs): NCCL winner check at 8 ranks, default bench line at N=8 (plot cut for the batch of stands, hypothesis cut for the single stand)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/r02_c32_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -1 gpurun_out/r02_c32_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c32_bench_n8.json 2> gpurun_out/r02_c32_bench_n8.err; echo "bench rc=$?"; tail -1 gpurun_out/r02_c32_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c32_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'resident', d['e2e_resident_index']['value']/1e6, d['config']['parallelism'])
print('single', d.get('single_stand'))
PY
