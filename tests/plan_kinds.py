"""Plan kinds added in round 2 -- general Rader (every "easy" prime), MixedRadix{r0 x Rader}, Good-Thomas, Bluestein over a
smooth inner length, caller-owned recipes -- as ONE list of cases run twice: on the CPU replay (tests/test_emu_parity.py) and on
the B200 (tests/test_gpu_parity.py), through the same C ABI.  The ranges follow the reference's own unit tests:
  raders_algorithm.rs:301-322     every prime in 3..100 with an explicit inner FFT; the 32-bit-overflow primes
  good_thomas_algorithm.rs:528-569 every coprime width x height in 1..12 / the butterfly sizes
  plan.rs:700-830                  recipe shapes (here: describe())"""
from math import gcd

import numpy as np
import pytest

import oracle
import rustfft_b200 as rb
from protocol import check_fft_algorithm
from rustfft_b200 import Recipe as R

DIRS = [rb.FftDirection.Forward, rb.FftDirection.Inverse]


def _is_prime(n):
    return n > 1 and all(n % d for d in range(2, int(n ** 0.5) + 1))


def _smooth(n):
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31):
        while n % p == 0:
            n //= p
    return n == 1


def check_rader_primes_below_100(planner, dtype):
    """raders_algorithm.rs:301-308 -- RadersAlgorithm::new(inner) for every prime 3..100, both directions.  The inner FFT here must be
    smooth (prime factors <= 31): 83 (82 = 2 x 41) is rejected by the recipe and planned as Bluestein by the library."""
    for p in range(3, 100):
        if not _is_prime(p):
            continue
        for d in DIRS:
            if _smooth(p - 1):
                f = check_fft_algorithm(planner, p, d, dtype, recipe=R.rader(p))
                assert f.describe().startswith("Rader{n=%d," % p), f.describe()
            else:
                with pytest.raises(rb.FftError, match="RADER"):
                    planner.plan_fft_with_recipe(R.rader(p), d)
                assert check_fft_algorithm(planner, p, d, dtype).describe().startswith("Bluestein{")


def check_default_prime_rule(planner, dtype):
    """src/plan.rs:636-664: Rader when p - 1 factors into small primes, Bluestein otherwise -- re-decided by measurement on the B200
    (profiles/r2d_ab_plans.txt): the one-pass Rader is the default in f64 and, in f32, from the size where Bluestein needs M >= 2048;
    above the one-pass limit the four run-time-radix passes lose to Bluestein over a power of two and stay recipe-only."""
    f32 = dtype == np.complex64
    cases = [(617, "Rader{n=617,g=3,inner=Smooth{616=11x7x8},fused}"), (719, "Bluestein{"), (2053, "Rader{n=2053,g=2,inner=Smooth{2052=19x3x3x3x4},fused}"),
             (1009, "Rader{n=1009,g=11,inner=Smooth{1008=7x3x3x16},fused}"), (257, "Rader{n=257,g=3,fused}"),
             (97, "Bluestein{n=97,M=256,fused}" if f32 else "Rader{n=97,g=5,inner=Smooth{96=3x16x2},fused}"),
             (4051, "Rader{n=4051,g=10,inner=Smooth{4050=5x5x3x3x3x3x2},fused}" if f32 else "Bluestein{n=4051,M=8192,inner=FourStep{64x128}}"),
             (7681, "Bluestein{n=7681,M=16384,inner=FourStep{128x128}}")]
    for n, want in cases:
        f = check_fft_algorithm(planner, n, DIRS[n % 2], dtype, control_kind=oracle.PLANNER, chunks=40 if n < 3000 else 3)
        assert f.describe().startswith(want), (n, f.describe())
    for n, want in [(4051, "Rader{n=4051,g=10,inner=SmoothFourStep{54x75}}" if not f32 else None), (7681, "Rader{n=7681,g=17,inner=SmoothFourStep{80x96}}")]:
        if want:
            f = check_fft_algorithm(planner, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=3, recipe=R.rader(n))
            assert f.describe() == want, f.describe()


def check_mixed_radix_rader(planner, dtype):
    """len = r0 x p, p an easy prime: the reference's MixedRadix{r0, Rader(p)} (1234 = 2 x 617 is BASELINE config 3's plan),
    here one CTA pass.  More virtual transforms than one CTA holds, ragged last CTA.  (Small lengths are planned as Bluestein by default
    in f32 -- measured faster -- so every case is also built from its recipe.)"""
    for n, r0, want in [(1234, 2, "MixedRadix{2xRader{n=617,g=3,inner=Smooth{616=11x7x8},fused},fused}"), (94, 2, "MixedRadix{2xRader{n=47,"),
                        (188, 4, "MixedRadix{4xRader{n=47,"), (296, 8, "MixedRadix{8xRader{n=37,"), (606, 6, "MixedRadix{6xRader{n=101,"),
                        (2049, 3, "MixedRadix{3xRader{n=683,"), (335, 5, "MixedRadix{5xRader{n=67,"), (7 * 103, 7, "MixedRadix{7xRader{n=103,")]:
        for d in DIRS:
            f = check_fft_algorithm(planner, n, d, dtype, control_kind=oracle.PLANNER, chunks=37 if d == DIRS[0] else 2, recipe=R.rader(n, r0))
            assert f.describe().startswith(want), (n, f.describe())
    f = check_fft_algorithm(planner, 1234, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=5)  # the default plan of config 3's length
    assert f.describe().startswith("MixedRadix{2xRader{n=617,"), f.describe()


def check_overflow_primes(planner, dtype):
    """raders_algorithm.rs:311-322: 112501 (112500 = 2^2 3^2 5^5: Rader over a two-pass smooth inner FFT), 216569 and 417623
    (p - 1 has a factor above 31: Bluestein); index products exceed 32 bits in all three."""
    for n in (112501, 216569, 417623):  # default: Bluestein over a power of two (measured faster than Rader's four run-time-radix passes)
        f = check_fft_algorithm(planner, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=2)
        assert f.describe().startswith("Bluestein{n=%d," % n), f.describe()
    for d in DIRS:  # RadersAlgorithm::new(inner) for the first of them, as the reference's test builds it
        f = check_fft_algorithm(planner, 112501, d, dtype, control_kind=oracle.PLANNER, chunks=2, recipe=R.rader(112501))
        assert f.describe() == "Rader{n=112501,g=10,inner=SmoothFourStep{300x375}}", f.describe()
    with pytest.raises(rb.FftError, match="RADER"):
        planner.plan_fft_with_recipe(R.rader(216569), DIRS[0])


def check_good_thomas_small_pairs(planner, dtype):
    """good_thomas_algorithm.rs:528-549: every coprime width x height below 12 (width 1 is the identity map: not a two-pass plan)."""
    for w in range(2, 12):
        for h in range(w + 1, 12):
            if gcd(w, h) != 1:
                continue
            for d in DIRS:
                f = check_fft_algorithm(planner, w * h, d, dtype, recipe=R.good_thomas(w, h), chunks=5)
                assert f.describe() == "GoodThomas{%dx%d}" % (w, h)


def check_good_thomas_large(planner, dtype):
    """Coprime splits of real sizes (44100 = 196 x 225, 10000 = 16 x 625, 48000 = 128 x 375, 5005 = 65 x 77, 1200 = 25 x 48 -- the
    example of good_thomas_algorithm.rs:22-38) against the same length planned as SmoothFourStep: identical maths, other index maps."""
    for a, b in [(196, 225), (16, 625), (128, 375), (65, 77), (25, 48), (27, 1000)]:
        n = a * b
        for d in DIRS:
            f = check_fft_algorithm(planner, n, d, dtype, control_kind=oracle.PLANNER, recipe=R.good_thomas(a, b), chunks=3)
            assert f.describe() == "GoodThomas{%dx%d}" % (min(a, b), max(a, b))


def check_bluestein_inner_lengths(planner, dtype):
    """The inner FFT of Bluestein may be any length >= 2n - 1 (src/plan.rs:649-657 picks 3/4 of a power of two when it fits;
    src/avx/avx_planner.rs:945-994 searches 2^a 3^b): power of two, smooth one-pass, smooth two-pass."""
    for rc, want in [(R.bluestein(1234, R.smooth(2500)), "Bluestein{n=1234,M=2500,inner=Smooth{2500=5x5x5x5x4},fused}"),
                     (R.bluestein(1234, R.pow2(4096)), "Bluestein{n=1234,M=4096,fused}"),
                     (R.bluestein(1234, R.smooth(3072)), "Bluestein{n=1234,M=3072,inner=Smooth{3072=3x16x16x4},fused}"),  # the reference's own choice
                     (R.bluestein(719), "Bluestein{n=719,M=2048,fused}"),
                     (R.bluestein(4099, R.mixed_radix(84, 98)), "Bluestein{n=4099,M=8232,inner=SmoothFourStep{84x98}}"),
                     (R.bluestein(4099, R.pow2(16384)), "Bluestein{n=4099,M=16384,inner=FourStep{128x128}}")]:
        if dtype == np.complex128 and "3072" in want:
            continue
        for d in DIRS:
            f = check_fft_algorithm(planner, rc.len, d, dtype, control_kind=oracle.PLANNER, recipe=rc, chunks=3)
            assert f.describe() == want, f.describe()
    f = check_fft_algorithm(planner, 1283, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=9)  # default: the next power of two (measured faster)
    assert f.describe() == "Bluestein{n=1283,M=4096,fused}"
    for d in DIRS:
        f = check_fft_algorithm(planner, 1283, d, dtype, control_kind=oracle.PLANNER, chunks=9, recipe=R.bluestein(1283, R.smooth(2592)))
        assert f.describe() == "Bluestein{n=1283,M=2592,inner=Smooth{2592=3x3x3x3x16x2},fused}"


def check_recipes_of_existing_kinds(planner, dtype):
    for rc, want in [(R.pow2(1024), "Direct{1024}"), (R.pow2(1 << 15), "FourStep{128x256"), (R.mixed_radix(128, 256), "FourStep{128x256"),
                     (R.smooth(1000), "Smooth{1000=5x5x5x8}"), (R.mixed_radix(100, 100), "SmoothFourStep{100x100"), (R.mixed_radix(50, 100), "SmoothFourStep{50x100}"),
                     (R.rader(257), "Rader{n=257,g=3,fused}"), (R.rader(65537, 1, R.pow2(65536)), "Rader{n=65537,g=3,inner=FourStep{256x256}}"),
                     (R.bluestein(37), "Bluestein{n=37,M=128,fused}")]:
        f = check_fft_algorithm(planner, rc.len, DIRS[0], dtype, control_kind=oracle.PLANNER, recipe=rc, chunks=2)
        assert f.describe().startswith(want), f.describe()
    for bad, msg in [(R.good_thomas(100, 100), "coprime"), (R.mixed_radix(3, 5000), "smooth one-pass"), (R.rader(1000), "prime"),
                     (R.bluestein(100, R.pow2(128)), "2 len - 1"), (R.pow2(1000), "power-of-two"), (R.smooth(1234), "prime factors"),
                     (R.rader(83), "primes <= 31")]:
        with pytest.raises(rb.FftError, match=msg):
            planner.plan_fft_with_recipe(bad, DIRS[0])


def check_cluster_plans(planner, dtype):
    """f32 2^14 .. 2^17 in ONE pass over HBM: both four-step passes inside a thread-block cluster of 2 / 4 / 8 / 16 CTAs, the transpose
    between them through distributed shared memory (cluster.h); against the scalar-planner oracle (Radix4, src/algorithm/radix4.rs) and the
    chunked / fused two-pass plan of the same length (same butterflies and tables: results within rounding of each other)."""
    if dtype != np.complex64:
        with pytest.raises(rb.FftError, match="CLUSTER"):
            planner.plan_fft_with_recipe(R.cluster(1 << 15), DIRS[0])
        return
    from util import rel_l2, signal

    for lg, c in [(14, 2), (15, 4), (16, 8), (17, 16)]:
        n = 1 << lg
        for d in DIRS:
            f = check_fft_algorithm(planner, n, d, dtype, control_kind=oracle.PLANNER, chunks=5 if d == DIRS[0] else 1, recipe=R.cluster(n))
            assert f.describe() == "ClusterFourStep{%dx%d,cluster=%d}" % (1 << (lg // 2), 1 << (lg - lg // 2), c), f.describe()
        x = signal(3 * n, dtype, seed=lg)
        a, b = x.copy(), x.copy()
        planner.plan_fft_with_recipe(R.cluster(n), DIRS[0]).process(a)
        planner.plan_fft_forward(n).process(b)
        assert rel_l2(a, b) < 1e-6
    with pytest.raises(rb.FftError, match="CLUSTER"):
        planner.plan_fft_with_recipe(R.cluster(1 << 18), DIRS[0])
    for lg, c in [(14, 4), (15, 8), (16, 16)]:  # half tiles: 4096 points and 128 threads per CTA, twice the cluster size
        n = 1 << lg
        for d in DIRS:
            f = check_fft_algorithm(planner, n, d, dtype, control_kind=oracle.PLANNER, chunks=3, recipe=R.cluster(n, half_tiles=True))
            assert f.describe() == "ClusterFourStep{%dx%d,cluster=%d}" % (1 << (lg // 2), 1 << (lg - lg // 2), c), f.describe()
    # the whole Rader / Bluestein algorithm inside one cluster pass (BASELINE config 4: n = 65537)
    for rc, want in [(R.rader(65537, 1, R.cluster(65536)), "Rader{n=65537,g=3,inner=ClusterFourStep{256x256,cluster=8},fused}"),
                     (R.bluestein(20011, R.cluster(65536)), "Bluestein{n=20011,M=65536,inner=ClusterFourStep{256x256,cluster=8},fused}"),
                     (R.bluestein(32768, R.cluster(65536)), "Bluestein{n=32768,M=65536,inner=ClusterFourStep{256x256,cluster=8},fused}"),
                     (R.bluestein(6007, R.cluster(16384)), "Bluestein{n=6007,M=16384,inner=ClusterFourStep{128x128,cluster=2},fused}")]:
        for d in DIRS:
            f = check_fft_algorithm(planner, rc.len, d, dtype, control_kind=oracle.PLANNER, chunks=3 if d == DIRS[0] else 1, recipe=rc)
            assert f.describe() == want, f.describe()


def check_plan_serialisation(planner, dtype):
    """b200fft_plan_recipe: the decomposition a plan was built as comes back as data (JSON-able), and b200fft_plan_create_from_recipe rebuilds
    the same plan from it -- same description, bit-identical results (plan caching across processes)."""
    import json

    from util import signal

    for n in [8, 1024, 1 << 15, 1000, 5000, 10000, 44100, 617, 1234, 97, 257, 65537, 719, 4099]:
        for d in DIRS:
            f = planner.plan_fft(n, d)
            rc = R.from_dict(json.loads(json.dumps(f.recipe().to_dict())))
            g = planner.plan_fft_with_recipe(rc, d)
            assert g.describe() == f.describe() and g.len() == n, (n, f.describe(), g.describe())
            x = signal(3 * n, dtype, seed=n)
            a, b = x.copy(), x.copy()
            f.process(a)
            g.process(b)
            assert np.array_equal(a, b), n


ALL = [check_plan_serialisation, check_cluster_plans, check_rader_primes_below_100, check_default_prime_rule, check_mixed_radix_rader, check_overflow_primes, check_good_thomas_small_pairs,
       check_good_thomas_large, check_bluestein_inner_lengths, check_recipes_of_existing_kinds]
