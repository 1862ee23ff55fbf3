#!/bin/bash
set -x
OUT=gpurun_out/r1o
mkdir -p $OUT
for mb in 32 48; do
B200FFT_CHUNK_MB=$mb timeout 300 python bench.py --steps 3 --logs 16,18,20 --no-e2e --no-cpu > $OUT/bench_chunk$mb.json 2>> $OUT/bench.err
done
B200FFT_HOST_PIPE=3 timeout 600 python - > $OUT/hostpipe.txt 2>&1 <<'PY'
import numpy as np, torch, time, sys, os
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
f = pl.plan_fft_forward(4096)
n = 1 << 29  # 4 GiB of complex64
hin = torch.empty(n, dtype=torch.complex64).pin_memory(); hout = torch.empty_like(hin).pin_memory()
torch.view_as_real(hin).uniform_(0, 10)
a, b = hin.numpy(), hout.numpy()
for rep in range(3):
    t0 = time.perf_counter(); f.process_outofplace_with_scratch(a, b); t1 = time.perf_counter()
    print("3-stage: 4 GiB in + 4 GiB out in %.3f s -> %.1f GB/s per direction" % (t1 - t0, 4.295 / (t1 - t0)))
d = torch.empty(n, dtype=torch.complex64, device='cuda')
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hin, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("plain H2D 4 GiB: %.1f GB/s" % (4.295 / (t1 - t0)))
    t0 = time.perf_counter(); hout.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("plain D2H 4 GiB: %.1f GB/s" % (4.295 / (t1 - t0)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hin, non_blocking=True)
with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("concurrent H2D + D2H of 4 GiB each: %.3f s -> %.1f GB/s per direction" % (t1 - t0, 4.295 / (t1 - t0)))
PY
timeout 300 python - > $OUT/hostpipe2.txt 2>&1 <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
f = pl.plan_fft_forward(4096)
n = 1 << 29
hin = torch.empty(n, dtype=torch.complex64).pin_memory(); hout = torch.empty_like(hin).pin_memory()
torch.view_as_real(hin).uniform_(0, 10)
a, b = hin.numpy(), hout.numpy()
for rep in range(3):
    t0 = time.perf_counter(); f.process_outofplace_with_scratch(a, b); t1 = time.perf_counter()
    print("2-stream: 4 GiB in + 4 GiB out in %.3f s -> %.1f GB/s per direction" % (t1 - t0, 4.295 / (t1 - t0)))
PY
cat $OUT/hostpipe.txt $OUT/hostpipe2.txt
