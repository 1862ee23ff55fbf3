#!/bin/bash
set -x
OUT=gpurun_out/r1r
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python - > $OUT/timing.txt 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
def t(n, batch):
    f = pl.plan_fft_forward(n)
    x = torch.view_as_complex(torch.rand(batch*n, 2, device='cuda')).contiguous(); y = torch.empty_like(x)
    for _ in range(2): f.process_device(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f.process_device(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    print(f"n={n:9d} batch={batch:6d} {f.describe():50s} {ms:9.4f} ms  {16.0*n*batch/ms/1e6:8.1f} GB/s  frac={16.0*n*batch/ms/1e6/6487.4:.3f}", flush=True)
for n in [143, 961, 1196, 1131, 2014, 3003]:
    t(n, (1<<28)//n//8)
for lg in [21, 22, 23, 24]:
    t(1 << lg, (1 << 31) >> lg)
t(1000003, 64)
PY
cat $OUT/timing.txt
