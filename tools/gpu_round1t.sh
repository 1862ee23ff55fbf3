#!/bin/bash
# round-1 run t: dataflow four-step v2 (pipelined tickets, deferred release) + few-twiddle-load builds, A/B
set -x
OUT=gpurun_out/r1t
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
L=$PWD/rustfft_b200
timeout 300 python tests/variant_check.py > $OUT/variant_flow.log 2>&1; echo "rc=$?" >> $OUT/variant_flow.log; tail -2 $OUT/variant_flow.log
B200FFT_LIB=$L/libb200fft_few2.so timeout 300 python tests/variant_check.py > $OUT/variant_few2.log 2>&1; echo "rc=$?" >> $OUT/variant_few2.log; tail -2 $OUT/variant_few2.log
ALL=10,11,12,13,14,15,16,17,18,19,20
env timeout 200 python tools/ab_two_pass.py $ALL >> $OUT/ab.log 2>&1
env B200FFT_FLOW=0 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_LIB=$L/libb200fft_few.so timeout 200 python tools/ab_two_pass.py $ALL >> $OUT/ab.log 2>&1
env B200FFT_LIB=$L/libb200fft_few2.so timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_LIB=$L/libb200fft_few2.so B200FFT_FLOW=0 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_FLOW_LOOKAHEAD=300 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_FLOW_LOOKAHEAD=900 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
grep SUMMARY $OUT/ab.log
