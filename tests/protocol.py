"""The reference's test protocol, restated once and run twice: on the CPU replay of the kernels
(tests/test_emu_parity.py, `-m "not gpu"`) and on the real CUDA path (tests/test_gpu_parity.py,
`-m gpu`), both through the same C ABI and the same FftPlanner / Fft mirror.

  check_fft_algorithm   src/test_utils.rs:70-209   four entry points x dirty scratch x 3 chunks vs Dft
  fft_matches_control   tests/accuracy.rs:39-82    planner vs Bluestein-over-Radix4 control
"""
import numpy as np
import pytest

import oracle
import rustfft_b200 as rb
from util import mean_abs_err, rel_l2, signal, strict_bound, truth


def dirty(n, dtype):
    # scratch filled with 100+100i, as the reference does (src/test_utils.rs:131-141)
    return np.full(n, 100 + 100j, dtype=dtype)


def check_fft_algorithm(planner, n, direction, dtype, control_kind=oracle.CONTROL, chunks=3, seed=None,
                        strict_factor=4.0, recipe=None):
    """All four entry points must agree with the oracle control under the reference's criterion
    (mean |a-b| < 0.1) AND the strict bound of SURVEY.md 8(c): relative L2 vs the f64 truth <= 4 eps log2 N -- always --
    AND stay near the oracle's own accuracy: not worse than 2x the oracle's error on the same input, unless the error is
    already below a quarter of the strict bound (the oracle is sometimes exceptionally accurate -- a Radix4 with f64-rounded
    twiddles against a Bluestein here -- and the measured worst ratio over len 1..1000 is 1.8 at 0.09 of the bound)."""
    inverse = direction == rb.FftDirection.Inverse
    # recipe: the caller owns planning (b200fft_plan_create_from_recipe) -- the way the reference's unit tests build one
    # algorithm directly (e.g. RadersAlgorithm::new(inner), src/algorithm/raders_algorithm.rs:324-329)
    fft = planner.plan_fft(n, direction) if recipe is None else planner.plan_fft_with_recipe(recipe, direction)
    assert fft.len() == n and fft.fft_direction() == direction
    x = signal(chunks * n, dtype, seed=n if seed is None else seed)
    want = oracle.fft(x, n, inverse, kind=control_kind)
    ref = truth(x, n, inverse)

    a = x.copy()
    fft.process(a)
    b = x.copy()
    fft.process_with_scratch(b, dirty(fft.get_inplace_scratch_len() + 7, dtype))
    c_in, c = x.copy(), dirty(x.size, dtype)
    fft.process_outofplace_with_scratch(c_in, c, dirty(fft.get_outofplace_scratch_len() + 3, dtype))
    d_in, d = x.copy(), dirty(x.size, dtype)
    d_in.setflags(write=False)
    fft.process_immutable_with_scratch(d_in, d, dirty(fft.get_immutable_scratch_len(), dtype))
    assert np.array_equal(d_in, x)

    for name, got in (("process", a), ("process_with_scratch", b), ("outofplace", c), ("immutable", d)):
        assert np.all(np.isfinite(got.view(got.real.dtype))), (n, name)
        assert mean_abs_err(got, want) < 0.1, (n, direction, name)
        assert np.array_equal(got, a), (n, name, "entry points disagree")
    if n >= 1:
        err = rel_l2(a, ref)
        oerr = rel_l2(want, ref)
        bound = strict_bound(n, dtype, strict_factor)
        assert err <= bound, (n, direction, err, bound, fft.describe())
        assert err <= 2.0 * oerr or err <= 0.25 * bound, (n, direction, err, oerr, fft.describe())
    return fft


def check_error_behaviour(planner, dtype):
    """The reference panics (src/common.rs:13-104); the mirror raises FftError with the same text."""
    fft = planner.plan_fft_forward(16)
    with pytest.raises(rb.FftError, match="Provided FFT buffer was too small. Expected len = 16, got len = 10"):
        fft.process(np.zeros(10, dtype))
    with pytest.raises(rb.FftError, match="Input FFT buffer must be a multiple of FFT length. Expected multiple of 16, got len = 40"):
        fft.process(np.zeros(40, dtype))
    with pytest.raises(rb.FftError, match="input buffer and output buffer must have the same length"):
        fft.process_outofplace_with_scratch(np.zeros(32, dtype), np.zeros(16, dtype))
    with pytest.raises(rb.FftError, match="input buffer and output buffer must have the same length"):
        fft.process_immutable_with_scratch(np.zeros(16, dtype), np.zeros(32, dtype))
    with pytest.raises(TypeError):
        fft.process(np.zeros(16, np.complex128 if dtype == np.complex64 else np.complex64))
    # len 0: silent no-op (src/fft_helper.rs:16-18); planning 0 and 1 must work (src/plan.rs:873-882)
    z = planner.plan_fft_forward(0)
    buf = signal(5, dtype)
    keep = buf.copy()
    z.process(buf)
    assert np.array_equal(buf, keep)
    one = planner.plan_fft_forward(1)
    one.process(buf)
    assert np.array_equal(buf, keep)
    out = np.zeros_like(buf)
    one.process_outofplace_with_scratch(buf.copy(), out)
    assert np.array_equal(out, keep)
    empty = np.zeros(0, dtype)
    fft.process(empty)  # zero chunks


def check_planner_cache(planner):
    # src/plan.rs:833-858: same (len, direction) -> same instance; other direction -> another
    a = planner.plan_fft(1234, rb.FftDirection.Forward)
    assert planner.plan_fft(1234, rb.FftDirection.Forward) is a
    assert planner.plan_fft_forward(1234) is a
    b = planner.plan_fft(1234, rb.FftDirection.Inverse)
    assert b is not a and planner.plan_fft_inverse(1234) is b
    assert a.fft_direction() == rb.FftDirection.Forward and b.fft_direction() == rb.FftDirection.Inverse
    assert a.fft_direction().opposite_direction() == rb.FftDirection.Inverse
    for f in (a, b):
        assert f.get_inplace_scratch_len() == 0 and f.get_outofplace_scratch_len() == 0 and f.get_immutable_scratch_len() == 0
