#!/bin/bash
set -x
OUT=gpurun_out/r1w
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
B200FFT_FLOW=1 timeout 300 python tests/variant_check.py > $OUT/variant_flow.log 2>&1; echo "rc=$?" >> $OUT/variant_flow.log; tail -2 $OUT/variant_flow.log
B200FFT_LIB=$PWD/rustfft_b200/libb200fft_trace.so B200FFT_FLOW=1 timeout 300 python tools/flow_trace.py > $OUT/flow_trace.txt 2>&1
cat $OUT/flow_trace.txt
for v in "B200FFT_FLOW=1" "B200FFT_FLOW=1 B200FFT_FLOW_LOOKAHEAD=500" "B200FFT_FLOW=1 B200FFT_FLOW_LOOKAHEAD=1000" "B200FFT_FLOW=1 B200FFT_FLOW_LOOKAHEAD=1400"; do
  env $v timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
