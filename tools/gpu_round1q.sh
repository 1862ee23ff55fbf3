#!/bin/bash
set -x
OUT=gpurun_out/r1q
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
    print("%s N=%d %s: %.1f GB/s per direction" % (os.environ.get("B200FFT_HOST_PIPE", "3"), N, f.describe(), 2 * m * 8 / 1e9 / (t1 - t0)))
PY
python /tmp/hp.py > $OUT/hp3.txt 2>&1
cat $OUT/hp3.txt
timeout 600 python bench.py --steps 3 --no-cpu --no-extras > $OUT/bench.json 2> $OUT/bench.err
