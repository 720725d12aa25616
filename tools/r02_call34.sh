#!/bin/bash
# GPU call 34 (1 GPU): CTA-per-ICP kernel with the nine fit sums on seven warps (default) vs on warp 0 (nosplit): probe + bit identity, team parity tests
mkdir -p gpurun_out
for v in b200 nosplit; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so timeout 200 python tools/strong_scaling_probe.py --worlds 1,4,8 --kernels warp,cta --reps 9 > gpurun_out/r02_c34_probe_$v.jsonl 2> gpurun_out/r02_c34_probe.err
  python - $v <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_c34_probe_{sys.argv[1]}.jsonl"):
    d = json.loads(l); print(sys.argv[1], d["world"], d["kernel"], round(d["ms_median"], 4), round(d["ms_min"], 4), d["bit_identical_to_w1_warp"], d["passes"], d["searched"])
PY
  tail -1 gpurun_out/r02_c34_probe.err
done
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/r02_c34_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c34_pytest.log; tail -3 gpurun_out/r02_c34_pytest.log
timeout 100 python tools/fuzz_parity.py 60 77 > gpurun_out/r02_c34_fuzz.log 2>&1; tail -2 gpurun_out/r02_c34_fuzz.log
