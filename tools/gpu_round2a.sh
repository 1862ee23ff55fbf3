#!/bin/bash
# TMA tiles with table prefetch, 2 vs 3 resident CTAs, 2..4 streams
set -x
OUT=gpurun_out/r2a
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
L=$PWD/rustfft_b200
B200FFT_TMA_TILES=1 timeout 300 python tests/variant_check.py > $OUT/variant_tma.log 2>&1; tail -1 $OUT/variant_tma.log
B200FFT_LIB=$L/libb200fft_minb3.so B200FFT_TMA_TILES=1 timeout 300 python tests/variant_check.py > $OUT/variant_tma3.log 2>&1; tail -1 $OUT/variant_tma3.log
for v in "B200FFT_STREAMS=3" "B200FFT_TMA_TILES=1" "B200FFT_TMA_TILES=1 B200FFT_STREAMS=3" "B200FFT_TMA_TILES=1 B200FFT_STREAMS=4" "B200FFT_TMA_TILES=1 B200FFT_STREAMS=3 B200FFT_CHUNK_MB=96" "B200FFT_LIB=$L/libb200fft_minb3.so B200FFT_TMA_TILES=1" "B200FFT_LIB=$L/libb200fft_minb3.so B200FFT_TMA_TILES=1 B200FFT_STREAMS=3" "B200FFT_LIB=$L/libb200fft_minb3.so B200FFT_TMA_TILES=1 B200FFT_STREAMS=3 B200FFT_CHUNK_MB=96"; do
  env $v timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log | sed 's|/tmp/code/ejmahler__RustFFT/repo/rustfft_b200/||'
