#!/bin/bash
# GPU call 37 (1 GPU): CTA-per-ICP kernel A/B of two builds (VARIANTS="a b a b"): probe + bit identity
mkdir -p gpurun_out
for v in ${VARIANTS:-b200 oldsort b200 oldsort}; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so timeout 200 python tools/strong_scaling_probe.py --worlds 1,4,8 --kernels warp,cta --reps 9 > gpurun_out/r02_c37_probe_$v.jsonl 2> gpurun_out/r02_c37_probe.err
  python - $v <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_c37_probe_{sys.argv[1]}.jsonl"):
    d = json.loads(l)
    if d["kernel"] == "cta": print(sys.argv[1], d["world"], d["kernel"], round(d["ms_median"], 4), round(d["ms_min"], 4), d["bit_identical_to_w1_warp"], d["passes"])
PY
  tail -1 gpurun_out/r02_c37_probe.err
done
