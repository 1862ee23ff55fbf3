#!/bin/bash
# session 9: ring size of the fused kernel at the two largest sizes (the 64 MiB cap binds there: W = 8 at 2^20, 16 at 2^19)
OUT=gpurun_out/s9
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
for w in 4 6 8 10 12; do
  B200FFT_FUSED_W=$w timeout 200 python tools/ab_two_pass.py 20 > $OUT/ab_w${w}_20.txt 2>&1; tail -1 $OUT/ab_w${w}_20.txt
done
for w in 8 12 16 20 24; do
  B200FFT_FUSED_W=$w timeout 200 python tools/ab_two_pass.py 19 > $OUT/ab_w${w}_19.txt 2>&1; tail -1 $OUT/ab_w${w}_19.txt
done
for w in 16 32 48; do
  B200FFT_FUSED_W=$w timeout 200 python tools/ab_two_pass.py 18 > $OUT/ab_w${w}_18.txt 2>&1; tail -1 $OUT/ab_w${w}_18.txt
done
