#!/bin/bash
# session 8: A/B of the fused kernel's direct pass-B output (B200FFT_FUSED_BDIRECT) per size, then its parity variant test
OUT=gpurun_out/s8
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_default.txt 2>&1; tail -1 $OUT/ab_default.txt
B200FFT_FUSED_BDIRECT=63 timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_bdirect.txt 2>&1; tail -1 $OUT/ab_bdirect.txt
timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_default2.txt 2>&1; tail -1 $OUT/ab_default2.txt
B200FFT_FUSED_BDIRECT=63 timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_bdirect2.txt 2>&1; tail -1 $OUT/ab_bdirect2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh_process and fused" > $OUT/pytest_variants.log 2>&1; tail -3 $OUT/pytest_variants.log
