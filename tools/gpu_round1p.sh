#!/bin/bash
set -x
OUT=gpurun_out/r1p
mkdir -p $OUT
cat > /tmp/hp.py <<'PY'
import numpy as np, torch, time, sys, os
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
n = 1 << 29
hin = torch.empty(n, dtype=torch.complex64).pin_memory(); hout = torch.empty_like(hin).pin_memory()
torch.view_as_real(hin).uniform_(0, 10)
a, b = hin.numpy(), hout.numpy()
for N in (1024, 4096, 1 << 16, 1 << 20, 1000, 65537):
    f = pl.plan_fft_forward(N)
    m = (n // N) * N
    f.process_outofplace_with_scratch(a[:m], b[:m])
    t0 = time.perf_counter()
    for _ in range(2): f.process_outofplace_with_scratch(a[:m], b[:m])
    t1 = time.perf_counter()
    print("%s N=%d %s: %.1f GB/s per direction" % (os.environ.get("B200FFT_HOST_PIPE", "2"), N, f.describe(), 2 * m * 8 / 1e9 / (t1 - t0)))
PY
B200FFT_HOST_PIPE=3 python /tmp/hp.py > $OUT/hp3.txt 2>&1
B200FFT_HOST_PIPE=2 python /tmp/hp.py > $OUT/hp2.txt 2>&1
cat $OUT/hp3.txt $OUT/hp2.txt
python - > $OUT/smooth_timing.txt 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
def t(n, batch):
    f = pl.plan_fft_forward(n)
    x = torch.view_as_complex(torch.rand(batch*n, 2, device='cuda')).contiguous(); y = torch.empty_like(x)
    for _ in range(3): f.process_device(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f.process_device(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print(f"n={n:6d} batch={batch} {f.describe():45s} {ms:8.4f} ms  {16.0*n*batch/ms/1e6:8.1f} GB/s  frac={16.0*n*batch/ms/1e6/6487.4:.3f}")
for n in [1000, 1200, 1536, 2000, 2401, 3000, 3600, 4000, 360, 100]:
    t(n, (1<<28)//n//8)
PY
cat $OUT/smooth_timing.txt
