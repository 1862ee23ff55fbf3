"""CPU (`-m "not gpu"`): size-independent properties of the transform, through the CPU replay of the kernels (tests/emu)
for one length of every plan kind -- linearity, the shift theorem, Parseval, forward/inverse round trip, and batch
independence (src/array_utils.rs:151-177: chunks of one buffer are independent transforms)."""
import numpy as np
import pytest

import rustfft_b200 as rb
from util import emu_library, rel_l2, signal

# Direct, Smooth, fused Rader, fused Bluestein, FourStep (TMA-tiled passes), SmoothFourStep, Bluestein over FourStep, Rader over FourStep
LENS = [1024, 3000, 257, 1234, 1 << 15, 10000, 5003, 65537]
EPS = {np.dtype(np.complex64): 5.96e-8, np.dtype(np.complex128): 1.11e-16}


@pytest.fixture(scope="module", params=[np.complex64, np.complex128], ids=["f32", "f64"])
def planner(request):
    return rb.FftPlanner(request.param, lib=emu_library()), request.param


@pytest.mark.parametrize("n", LENS)
def test_linearity_shift_parseval_roundtrip(planner, n):
    pl, dtype = planner
    tol = 64 * EPS[np.dtype(dtype)]
    f, fi = pl.plan_fft_forward(n), pl.plan_fft_inverse(n)
    a, b = signal(n, dtype, 1), signal(n, dtype, 2)
    fa, fb, fab = a.copy(), b.copy(), (a + 3 * b).astype(dtype)
    f.process(fa), f.process(fb), f.process(fab)
    assert rel_l2(fab, fa + 3 * fb) < tol, f.describe()
    # shift theorem: x[(m - s) mod n]  <->  X[k] * exp(-2 pi i k s / n)
    s = 7 % n
    sh = np.roll(a, s)
    f.process(sh)
    k = np.arange(n)
    assert rel_l2(sh, fa.astype(np.complex128) * np.exp(-2j * np.pi * k * s / n)) < tol, f.describe()
    # Parseval (unnormalised transform: sum |X|^2 = n sum |x|^2, src/lib.rs:81-85)
    e_t = np.sum(np.abs(a.astype(np.complex128)) ** 2)
    e_f = np.sum(np.abs(fa.astype(np.complex128)) ** 2)
    assert abs(e_f / n / e_t - 1) < tol
    back = fa.copy()
    fi.process(back)
    assert rel_l2(back / n, a) < tol


@pytest.mark.parametrize("n", LENS)
def test_chunks_of_a_buffer_are_independent_transforms(planner, n):
    pl, dtype = planner
    f = pl.plan_fft_forward(n)
    batch = 5
    x = signal(n * batch, dtype, seed=n)
    whole = x.copy()
    f.process(whole)
    for b in (0, 2, 4):
        one = x[b * n:(b + 1) * n].copy()
        f.process(one)
        assert np.array_equal(one, whole[b * n:(b + 1) * n]), (f.describe(), b)
