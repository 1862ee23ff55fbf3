"""A/B timing of the two-pass (FourStep) plans under the B200FFT_* switches of the environment: f32 forward,
N = 2^15..2^20, 8 GiB of signal per size (inputs larger than L2), CUDA events around 3 execs after 2 warm-ups."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfft_b200 as rb

HBM = 6487.4
logs = [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "15,16,17,18,19,20".split(","))]
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B200FFT_")) or "default"
pl = rb.FftPlanner(np.complex64)
total = 1 << int(os.environ.get('AB_TOTAL_LOG2', '30'))
x = torch.empty(total, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(0, 10)
y = torch.empty_like(x)
out = []
for lg in logs:
    n = 1 << lg
    batch = total // n
    f = pl.plan_fft_forward(n)
    ws = torch.empty(max(16, f.workspace_bytes(batch)), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        f.process_device(x, out=y, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f.process_device(x, out=y, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    frac = 16.0 * total / ms / 1e6 / HBM
    out.append(f"2^{lg}:{frac:.3f}")
    print(f"[{tag}] n=2^{lg} batch={batch} {f.describe():40s} {ms:8.3f} ms frac={frac:.3f}", flush=True)
print(f"SUMMARY [{tag}] " + " ".join(out), flush=True)
