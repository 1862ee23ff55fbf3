#!/bin/bash
set -x
OUT=gpurun_out/r1v
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
B200FFT_LIB=$PWD/rustfft_b200/libb200fft_trace.so B200FFT_FLOW=1 timeout 300 python tools/flow_trace.py > $OUT/flow_trace.txt 2>&1
cat $OUT/flow_trace.txt
