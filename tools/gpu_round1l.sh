#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
python - > gpurun_out/smooth_timing.txt 2>&1 <<'PY'
import numpy as np, torch, math, sys
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
for n in [1000, 1024, 1200, 1234, 1536, 2000, 2048, 2401, 3000, 3600, 4000, 4096, 997, 1009]:
    t(n, (1<<28)//n//8)
PY
ls -la gpurun_out
