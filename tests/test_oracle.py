"""Pins the CPU oracle (oracle/rustfft_scalar_oracle.cpp) against the reference's own
known-answer tests and an f64 numpy ground truth.  CPU only.

Reference tests restated here (paths relative to /root/reference):
  src/algorithm/dft.rs:283-398      known signal -> spectrum pairs, N = 2, 3, 4, 6
  src/math_utils.rs:495-540         modular_exponent / primitive_root / distinct_prime_factors
  src/math_utils.rs:617-700         PrimeFactors table + partition consistency
  src/plan.rs:700-830               recipe shape per size class (scalar planner)
  src/twiddles.rs:77-98             rotate_90 == multiply by twiddle(1, 4)
  src/test_utils.rs:36-43,70-209    compare_vectors criterion; 3 chunks per buffer
  tests/accuracy.rs:98-187          planner output vs Bluestein-over-Radix4 control, len 1..1000
"""
import numpy as np
import pytest

import oracle


def _signal(n, dtype, seed=0):
    # the reference's distribution: re, im ~ U[0, 10)  (tests/accuracy.rs:84-95)
    rng = np.random.default_rng(seed)
    x = (rng.random(n) + 1j * rng.random(n)) * 10
    return x.astype(dtype)


def _mean_abs_err(a, b):
    # src/test_utils.rs:36-43 / tests/accuracy.rs:30-37
    return float(np.mean(np.abs(a.astype(np.complex128) - b.astype(np.complex128))))


def _truth(x, n, inverse):
    x = x.astype(np.complex128).reshape(-1, n)
    return (np.fft.ifft(x, axis=1) * n if inverse else np.fft.fft(x, axis=1)).ravel()


DFT_KATS = [
    ([1, -1], [0, 2]),
    ([1 + 1j, 2 - 3j, -1 + 4j], [2 + 2j, -5.562177 - 2.098076j, 6.562178 + 3.09807j]),
    ([1j, 2.5 - 3j, -1 - 1j, 4], [5.5 - 3j, -2 + 3.5j, -7.5 + 3j, 4 + 0.5j]),
    ([1 + 1j, 2 + 2j, 3 + 3j, 4 + 4j, 5 + 5j, 6 + 6j],
     [21 + 21j, -8.16 + 2.16j, -4.76 - 1.24j, -3 - 3j, -1.24 - 4.76j, 2.16 - 8.16j]),
]


@pytest.mark.parametrize("kind", [oracle.DFT, oracle.PLANNER, oracle.CONTROL])
def test_dft_known_answers(kind):
    for sig, spec in DFT_KATS:
        x = np.array(sig, dtype=np.complex64)
        y = oracle.fft(x, len(sig), False, kind=kind)
        assert _mean_abs_err(y, np.array(spec)) < 0.1
        assert np.max(np.abs(y - np.array(spec))) < 0.06  # the len-6 literals are only good to ~0.05


def test_math_utils_known_answers():
    for (b, e, m), want in [((2, 8, 300), 256), ((2, 9, 300), 212), ((1, 9, 300), 1), ((3, 416788, 47), 8)]:
        assert oracle.modular_exponent(b, e, m) == want
    for p, want in [(3, 2), (7, 3), (11, 2), (13, 2), (47, 5), (7919, 7)]:
        assert oracle.primitive_root(p) == want
    for n, want in [(46, [2, 23]), (2, [2]), (3, [3]), (162, [2, 3])]:
        assert oracle.distinct_prime_factors(n) == want
    assert oracle.primitive_root(65537) == 3 and oracle.primitive_root(617) == 3


def test_prime_factors_table():
    table = [
        (2, {2: 1}, 1, 1, True), (128, {2: 7}, 7, 1, False), (3, {3: 1}, 1, 1, True),
        (81, {3: 4}, 4, 1, False), (5, {5: 1}, 1, 1, True), (125, {5: 3}, 3, 1, False),
        (97, {97: 1}, 1, 1, True), (6, {2: 1, 3: 1}, 2, 2, False), (12, {2: 2, 3: 1}, 3, 2, False),
        (36, {2: 2, 3: 2}, 4, 2, False), (10, {2: 1, 5: 1}, 2, 2, False),
        (100, {2: 2, 5: 2}, 4, 2, False), (44100, {2: 2, 3: 2, 5: 2, 7: 2}, 8, 4, False),
    ]
    for n, fac, total, distinct, prime in table:
        f = oracle.prime_factors(n)
        assert f["total"] == total and f["distinct"] == distinct
        assert (f["total"] == 1) == prime
        assert f["p2"] == fac.get(2, 0) and f["p3"] == fac.get(3, 0)
        assert dict(f["other"]) == {k: v for k, v in fac.items() if k > 3}
    for n in range(4, 200):
        if oracle.prime_factors(n)["total"] > 1:
            l, r = oracle.partition_factors(n)
            assert l > 1 and r > 1 and l * r == n


def test_planner_recipe_shapes():
    assert oracle.describe_plan(0) == "Dft(0)" and oracle.describe_plan(1) == "Dft(1)"
    for pw in range(6, 25):
        assert oracle.describe_plan(1 << pw).startswith("Radix4{")
    for n in [2, 3, 4, 5, 6, 7, 8, 11, 12, 13, 16, 17, 19, 23, 24, 29, 31, 32]:
        assert oracle.describe_plan(n) == f"Butterfly{n}"
    for p2 in range(2, 5):
        for p3 in range(2, 5):
            for p5 in range(2, 4):
                for p7 in range(2, 4):
                    assert oracle.describe_plan(2**p2 * 3**p3 * 5**p5 * 7**p7).startswith("RadixN{")
    for n in [12 * 3, 6 * 27]:
        assert oracle.describe_plan(n).startswith("MixedRadixSmall{")
    for n in [15, 21, 35, 143]:
        assert oracle.describe_plan(n).startswith("GoodThomasAlgorithmSmall{")
    for n in [59, 83, 107, 149, 167, 173, 179, 359, 719, 1439, 2879]:
        assert oracle.describe_plan(n).startswith("BluesteinsAlgorithm{")
    for n in [53, 61, 67, 71, 73, 79, 89, 97, 101, 103, 109, 113, 127, 131, 137, 139, 151, 157, 163, 181, 191,
              193, 197, 199]:
        assert oracle.describe_plan(n).startswith("RadersAlgorithm{")
    # the plans SURVEY.md 3.1 derives by hand for the BASELINE configs
    assert oracle.describe_plan(1024) == "Radix4{k=3,base=Butterfly16}"
    assert oracle.describe_plan(1 << 16) == "Radix4{k=6,base=Butterfly16}"
    assert oracle.describe_plan(1 << 17) == "Radix4{k=7,base=Butterfly8}"
    assert oracle.describe_plan(65537) == "RadersAlgorithm{Radix4{k=6,base=Butterfly16}}"
    assert oracle.describe_plan(1234) == "RadixN{[2],base=RadersAlgorithm{RadixN{[7,2,4],base=Butterfly11}}}"


def test_twiddles():
    for inv in (False, True):
        t = oracle.twiddle(1, 4, inv)
        v = 3.0 + 4.0j
        rot = complex(-v.imag, v.real) if inv else complex(v.imag, -v.real)
        assert abs(v * t - rot) < 1e-15  # src/twiddles.rs:77-98
    # f32 twiddles are the f64 value rounded once (src/twiddles.rs:11-17)
    for k, n in [(1, 7), (5, 1024), (777, 65536), (123456, 1 << 20)]:
        t64 = np.exp(-2j * np.pi * k / n)
        t32 = oracle.twiddle(k, n, False, np.complex64)
        assert t32.real == np.float32(t64.real) and t32.imag == np.float32(t64.imag)


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 4 * 5.96e-8), (np.complex128, 4 * 1.11e-16)])
def test_every_len_1_to_1000_matches_control_and_truth(dtype, tol):
    """tests/accuracy.rs:124-187 restated for the oracle itself: planner vs control, both
    directions, 3 chunks per buffer, the reference's criterion -- plus the strict relative-L2 bound
    of SURVEY.md 8(c) against an f64 truth."""
    worst = 0.0
    for n in range(1, 1001):
        x = _signal(3 * n, dtype, seed=n)
        for inv in (False, True):
            got = oracle.fft(x, n, inv, kind=oracle.PLANNER)
            ctl = oracle.fft(x, n, inv, kind=oracle.CONTROL)
            assert _mean_abs_err(got, ctl) < 0.1, (n, inv)
            ref = _truth(x, n, inv)
            rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            bound = tol * max(1.0, np.log2(max(n, 2)))
            assert rel <= bound, (n, inv, rel, bound, oracle.describe_plan(n))
            worst = max(worst, rel / bound)
    assert worst <= 1.0


@pytest.mark.parametrize("n", [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1234, 65537, 112501])
def test_baseline_config_sizes(n):
    # 112501: one of the reference's 32-bit-overflow Rader primes (raders_algorithm.rs:311-322)
    # f64 gets 16x: the reference rounds the *angle* -2*pi*k/N in f64 before cos/sin
    # (src/twiddles.rs:11-12), so its f64 twiddles are only good to a few eps, and Rader's on a
    # U[0,10) signal (large DC) amplifies that (measured 5.2 eps*log2 N at 65537); the f64 numpy
    # "truth" carries its own eps*log2 N as well.
    for dtype, tol in [(np.complex64, 4 * 5.96e-8), (np.complex128, 16 * 1.11e-16)]:
        x = _signal(2 * n, dtype, seed=7)
        for inv in (False, True):
            got = oracle.fft(x, n, inv)
            ref = _truth(x, n, inv)
            assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= tol * np.log2(n)


def test_batch_is_independent_chunks_and_threads_agree():
    n = 96
    x = _signal(8 * n, np.complex64, seed=3)
    whole = oracle.fft(x, n)
    for b in range(8):
        np.testing.assert_array_equal(whole[b * n:(b + 1) * n], oracle.fft(x[b * n:(b + 1) * n], n))
    np.testing.assert_array_equal(whole, oracle.fft(x, n, threads=4))
    assert oracle.fft(x, 0).tolist() == x.tolist()  # len 0 is a silent no-op (src/fft_helper.rs:16-18)
    with pytest.raises(ValueError):
        oracle.fft(x[:100], n)


def test_linearity_and_roundtrip():
    n = 1234
    a, b = _signal(n, np.complex128, 1), _signal(n, np.complex128, 2)
    fa, fb, fab = oracle.fft(a, n), oracle.fft(b, n), oracle.fft(a + 2 * b, n)
    assert np.linalg.norm(fab - (fa + 2 * fb)) / np.linalg.norm(fab) < 1e-14
    back = oracle.fft(fa, n, inverse=True) / n  # unnormalised (src/lib.rs:81-85)
    assert np.linalg.norm(back - a) / np.linalg.norm(a) < 1e-14
