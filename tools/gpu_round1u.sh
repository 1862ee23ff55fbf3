#!/bin/bash
# round-1 run u: TMA-tiled four-step passes (B200FFT_TMA_TILES=1) vs the LDG/STG passes, few-twiddle-loads build
set -x
OUT=gpurun_out/r1u
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
B200FFT_TMA_TILES=1 timeout 300 python tests/variant_check.py > $OUT/variant_tma.log 2>&1; echo "rc=$?" >> $OUT/variant_tma.log; tail -4 $OUT/variant_tma.log
if ! grep -q VARIANT-OK $OUT/variant_tma.log; then
  B200FFT_TMA_TILES=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tests/tma_small_check.py > $OUT/sanitizer_tma.log 2>&1
  grep -v "^=========     Host Frame\|^=========         in " $OUT/sanitizer_tma.log | head -60
  exit 0
fi
ALL=10,11,12,13,14,15,16,17,18,19,20
env B200FFT_TMA_TILES=1 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_TMA_TILES=1 B200FFT_OVERLAP=0 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_TMA_TILES=1 B200FFT_STREAMS=3 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
env B200FFT_TMA_TILES=1 B200FFT_CHUNK_MB=96 timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
grep SUMMARY $OUT/ab.log
# ncu of the two TMA passes of 2^20 (one chunk pair) if the parity check passed
if grep -q VARIANT-OK $OUT/variant_tma.log; then
B200FFT_TMA_TILES=1 timeout 300 ncu --set full --clock-control none -k regex:run_kernel_tma -s 4 -c 2 -o /tmp/full_tma python tools/ab_two_pass.py 20 > $OUT/ncu_tma.log 2>&1
python tools/ncu_summary.py /tmp/full_tma.ncu-rep > $OUT/ncu_full_tma_1024.md 2>&1
ncu -i /tmp/full_tma.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_full_tma_1024_raw.csv.gz
fi
ls -la $OUT
