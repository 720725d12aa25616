#!/bin/bash
# A/B helper (run on the GPU box): bench the default workload against several builds of the library.
#   make -C coregistrationgame_b200/csrc variant NAME=tf FLAGS=-DFICP_TIEFREE_STREAM     (here, cross-compiled)
#   gpurun -- 'bash tools/ab_variants.sh b200 tf'                                        (one bench line per build)
# Extra bench arguments go in BENCH_ARGS, e.g. BENCH_ARGS="--dims 2".  Results: gpurun_out/ab_<name>.json
B="python bench.py --no-cpu-baseline --no-e2e --no-single-stand --steps 5 --warmup 3 $BENCH_ARGS"
mkdir -p gpurun_out
for v in "$@"; do
  FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_$v.so $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"] / 1e6, 2), "M hyp-iter/s", round(d["ms_per_step"], 2), "ms/step", d["path_stats"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
