#!/bin/bash
# GPU call 20 (8 GPUs): where the free-running N=8 end-to-end step goes
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29561 tools/e2e_breakdown.py 2> gpurun_out/r02_c20_bd.err | tee gpurun_out/r02_c20_breakdown_n8.jsonl
tail -3 gpurun_out/r02_c20_bd.err | cut -c1-300
