"""Real-input / real-output transforms (RealFft, b200fft_real_*): one list of checks run on the CPU replay and on the B200.
Truth = numpy rfft / irfft in f64; tolerance = the complex path's strict bound (4 eps log2 N relative L2)."""
import numpy as np
import pytest

import rustfft_b200 as rb
from util import EPS, rel_l2

LENGTHS = [2, 4, 6, 10, 16, 30, 100, 256, 1000, 1024, 1234, 2048, 4098, 10000, 44100, 65536, 100000, 1 << 17]


def check_real_fft(planner, rdtype, lengths=LENGTHS):
    cdtype = np.complex64 if rdtype == np.float32 else np.complex128
    eps = EPS[np.dtype(cdtype)]
    rng = np.random.default_rng(7)
    for n in lengths:
        f = planner.plan_fft(n)
        assert f.len() == n and f.complex_len() == n // 2 + 1
        batch = 5 if n <= 4096 else 2
        x = (rng.random(n * batch) * 10).astype(rdtype)  # the reference's test distribution, real part only
        X = np.zeros(batch * (n // 2 + 1), dtype=cdtype)
        f.forward(x, X)
        want = np.fft.rfft(x.astype(np.float64).reshape(batch, n), axis=1).ravel()
        bound = 4 * eps * max(1.0, np.log2(n))
        assert rel_l2(X, want) <= bound, (n, rel_l2(X, want), bound)
        back = np.zeros_like(x)
        f.inverse(X, back)
        assert rel_l2(back / n, x) <= 2 * bound, (n, rel_l2(back / n, x))  # unnormalised both ways, like the complex path
        y = np.zeros_like(x)
        f.inverse(want.astype(cdtype), y)  # inverse alone, from the exact spectrum
        assert rel_l2(y / n, x) <= bound, n
    with pytest.raises(rb.FftError, match="even length"):
        planner.plan_fft(7)
    f = planner.plan_fft(16)
    with pytest.raises(rb.FftError, match="expected batch"):
        f.forward(np.zeros(32, rdtype), np.zeros(16, cdtype))
    with pytest.raises(TypeError):
        f.forward(np.zeros(32, np.complex64 if rdtype == np.float32 else np.complex128), np.zeros(18, cdtype))
