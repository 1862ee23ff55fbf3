"""2-D complex transforms (Fft2d, b200fft_plan2d_*): checks shared by the CPU replay and the B200.  Truth = numpy fft2 / ifft2 in f64."""
import numpy as np
import pytest

import rustfft_b200 as rb
from util import EPS, rel_l2, signal

SHAPES = [(2, 2), (4, 4), (1, 16), (8, 16), (3, 5), (30, 50), (64, 1024), (100, 100), (31, 37), (270, 480), (1024, 64), (1080, 1920)]


def check_fft2d(planner, dtype, shapes=SHAPES):
    eps = EPS[np.dtype(dtype)]
    for h, w in shapes:
        batch = 3 if h * w <= 1 << 16 else 1
        for d in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
            f = planner.plan_fft_2d(h, w, d)
            x = signal(batch * h * w, dtype, seed=h * w)
            y = x.copy()
            f.process(y)
            xs = x.astype(np.complex128).reshape(batch, h, w)
            want = (np.fft.fft2(xs) if d == rb.FftDirection.Forward else np.fft.ifft2(xs) * (h * w)).ravel()
            bound = 4 * eps * max(1.0, np.log2(h * w))
            assert rel_l2(y, want) <= bound, (h, w, d, rel_l2(y, want), bound)
    with pytest.raises(rb.FftError, match="COLUMNS"):
        planner.plan_fft_2d(37 * 41, 8)  # a column length with prime factors above 31 has no column pass
