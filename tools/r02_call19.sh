#!/bin/bash
# GPU call 19 (8 GPUs): where the N=8 end-to-end step goes, NCCL check with the 14-word record, C4 at N=8 after the host-prep work
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tests/dist_check.py > gpurun_out/r02_c19_distcheck.log 2>&1; echo "dist_check rc=$?"; tail -1 gpurun_out/r02_c19_distcheck.log
timeout 300 $TR --nproc-per-node 8 --master-port 29561 tools/e2e_breakdown.py 2> gpurun_out/r02_c19_bd.err | tee gpurun_out/r02_c19_breakdown_n8.json
timeout 400 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 > gpurun_out/r02_c19_c4_n8.json 2> gpurun_out/r02_c19_c4_n8.err; echo "c4 n=8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c19_c4_n8.json').read().strip().splitlines()[-1])
print('c4 n8 value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], 'res', d.get('e2e_resident_index'))
PY
