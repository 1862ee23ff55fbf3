#!/bin/bash
# One GPU session (gpurun): parity of the fused two-pass kernel, then A/B timings (tools/ab_two_pass.py) of the two-pass switches.
# Results under gpurun_out/s1/.
OUT=gpurun_out/s1
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/env.txt 2>&1
# 1. quick parity of the new path (full suite later in the session)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or up_to_2_24 or device_path or ragged or native_library" > $OUT/pytest_quick.log 2>&1
echo "pytest quick rc=$?" >> $OUT/pytest_quick.log
tail -5 $OUT/pytest_quick.log
# 2. A/B: fused (default) vs ceiling vs chunked, look-ahead sweep
for v in "" \
         "B200FFT_FUSED_NOCOMPUTE=1" \
         "B200FFT_FUSED=0" \
         "B200FFT_FUSED_LOOKAHEAD=300" \
         "B200FFT_FUSED_LOOKAHEAD=1200" \
         "B200FFT_DISCARD=0" \
         "B200FFT_TMA_L2PROMO=2" \
         "B200FFT_LIB=$PWD/rustfft_b200/libb200fft_ps5.so" \
         "B200FFT_FUSED=0 B200FFT_NARROW=1" \
         "B200FFT_FUSED=0 B200FFT_NARROW=1 B200FFT_TMA_L2PROMO=2" \
         "B200FFT_FUSED=0 B200FFT_LIB=$PWD/rustfft_b200/libb200fft_ps5.so" \
         "B200FFT_FUSED=0 B200FFT_PREFETCH=1" \
         "B200FFT_FUSED=0 B200FFT_RADIX32=0"; do
  env $v timeout 200 python tools/ab_two_pass.py 14,15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
env B200FFT_PERSIST=1 timeout 100 python tools/ab_two_pass.py 14 >> $OUT/ab.log 2>&1
grep "PERSIST" $OUT/ab.log | tail -2
