#!/bin/bash
# GPU call 30 (1 GPU): CTA-per-ICP kernel experiments - group width by all threads (gs), skip-test rounds unrolled (su), both (gssu)
mkdir -p gpurun_out
for v in b200 gs su gssu; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so timeout 200 python tools/strong_scaling_probe.py --worlds 1,4,8 --kernels cta --reps 9 > gpurun_out/r02_c30_probe_$v.jsonl 2> gpurun_out/r02_c30_probe.err
  python - $v <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_c30_probe_{sys.argv[1]}.jsonl"):
    d = json.loads(l); print(sys.argv[1], d["world"], d["kernel"], round(d["ms_median"], 4), round(d["ms_min"], 4), d["bit_identical_to_w1_warp"], d["searched"])
PY
  tail -1 gpurun_out/r02_c30_probe.err
done
