#!/bin/bash
# Cheap A/B runs queued for the next GPU session (each line ~10 s on the box; the whole script ~3 GPU-minutes).
# Results: gpurun_out/next/ab.log -- fractions of the measured HBM roofline per size (tools/ab_two_pass.py).
# Before `gpurun`, on the CPU side:  (cd rustfft_b200/csrc && make -j5 OUT=../libb200fft_ps5.so BUILD=build_ps5 EXTRA=-DB2_TILE1024_PS=5)
set -x
OUT=gpurun_out/next
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
for v in "" \
         "B200FFT_RADIX32=0" \
         "B200FFT_TMA_L2PROMO=2" \
         "B200FFT_TMA_L2PROMO=3" \
         "B200FFT_STREAMS=4 B200FFT_CHUNK_MB=48" \
         "B200FFT_STREAMS=4 B200FFT_CHUNK_MB=80" \
         "B200FFT_PERSIST=1" \
         "B200FFT_PREFETCH=1" \
         "B200FFT_PREFETCH=1 B200FFT_CHUNK_MB=32" \
         "B200FFT_NARROW=1" \
         "B200FFT_NARROW=1 B200FFT_TMA_L2PROMO=2" \
         "B200FFT_LIB=$PWD/rustfft_b200/libb200fft_ps5.so" \
         "B200FFT_FLOW=1 B200FFT_FLOW_LOOKAHEAD=1000"; do
  env $v timeout 200 python tools/ab_two_pass.py 14,15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
# throughput of the SmoothFourStep plans (never timed in round 1: the GPU budget ran out after their parity run)
python - > $OUT/smooth_four_step.txt 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
for n in [5000, 10000, 44100, 48000, 100000, 1000000]:
    batch = (1 << 28) // n
    f = pl.plan_fft_forward(n)
    x = torch.view_as_complex(torch.rand(batch * n, 2, device='cuda')).contiguous(); y = torch.empty_like(x)
    for _ in range(2): f.process_device(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f.process_device(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"n={n:8d} batch={batch:6d} {f.describe():34s} {ms:9.4f} ms  frac={16.0*n*batch/ms/1e6/6487.4:.3f}", flush=True)
PY
cat $OUT/smooth_four_step.txt
