#!/bin/bash
# GPU call 1 of round 2: parity of the default build (incl. the new per-pass trace tests), parity of the tie-free
# stream build, A/B of the two, pass-count histogram of the literal C3 stand, fresh ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_default.log
tail -5 gpurun_out/r02_pytest_default.log
FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_tf.so python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_tf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_tf.log
tail -5 gpurun_out/r02_pytest_tf.log
bash tools/ab_variants.sh b200 tf 2>&1 | tee gpurun_out/r02_ab.log
python tools/pass_hist.py > gpurun_out/r02_passhist.json 2> gpurun_out/r02_passhist.err; cat gpurun_out/r02_passhist.json
CMD="python bench.py --no-cpu-baseline --no-e2e --no-single-stand --steps 1 --warmup 3"
export FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_tf.so
$CMD > gpurun_out/plain_tf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_kernel -s 3 -c 1 -f -o gpurun_out/r02_icp_tf $CMD > gpurun_out/ncu_tf.log 2>&1
echo "ncu tf rc=$?"
unset FICP_B200_LIB
$CMD > gpurun_out/plain_def.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:icp_kernel -s 3 -c 1 -f -o gpurun_out/r02_icp_def $CMD > gpurun_out/ncu_def.log 2>&1
echo "ncu def rc=$?"
ls -la gpurun_out/*.ncu-rep
