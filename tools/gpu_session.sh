#!/bin/bash
OUT=gpurun_out/s11
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_power or up_to_2_24 or ragged or device_path" > $OUT/pytest_quick.log 2>&1
echo "pytest quick rc=$?" >> $OUT/pytest_quick.log
tail -n 3 $OUT/pytest_quick.log
for v in "" "B200FFT_FUSED_TW2=0"; do
  env $v timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
