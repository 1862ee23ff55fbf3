#!/bin/bash
# session 2: new plan kinds on the GPU (parity), plan A/B table, sanitizer runs
OUT=gpurun_out/s2
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round2 or large_non_power or smooth or config3 or config4 or native_library" > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_new.log; tail -5 $OUT/pytest_new.log
AB_PLANS_OUT=$OUT/ab_plans.json timeout 600 python tools/ab_plans.py > $OUT/ab_plans.txt 2>&1; echo "ab rc=$?"; cat $OUT/ab_plans.txt
bash tools/gpu_sanitize.sh $OUT/sanitize 120 2>&1 | tail -12
