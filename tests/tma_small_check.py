"""Tiny two-pass case for compute-sanitizer runs (tools/gpu_round1*.sh): f32 N = 2^15, 3 transforms."""
import sys

import numpy as np

import rustfft_b200 as rb
from util import rel_l2, signal, truth

n, batch = 1 << 15, 3
f = rb.FftPlanner(np.complex64).plan_fft_forward(n)
x = signal(n * batch, np.complex64, seed=1)
y = x.copy()
f.process(y)
print(f.describe(), rel_l2(y, truth(x, n, False)))
