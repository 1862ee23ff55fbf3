"""CPU (`-m "not gpu"`): the library's planner, tables and kernel phase functions, replayed on the
CPU by tests/emu behind the real C ABI, checked with the reference's protocol against the oracle.
This is host-logic coverage; the parity tests proper are tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle
import rustfft_b200 as rb
import plan_kinds
from protocol import check_error_behaviour, check_fft_algorithm, check_planner_cache
from util import emu_library, rel_l2, signal, truth

DIRS = [rb.FftDirection.Forward, rb.FftDirection.Inverse]


@pytest.fixture(scope="module")
def lib():
    return emu_library()


@pytest.fixture(scope="module", params=[np.complex64, np.complex128], ids=["f32", "f64"])
def planner(request, lib):
    return rb.FftPlanner(request.param, lib=lib), request.param


def test_every_len_1_to_200_both_directions(planner):
    pl, dtype = planner
    for n in range(1, 201):
        for d in DIRS:
            check_fft_algorithm(pl, n, d, dtype)


def test_sampled_lens_up_to_1000_and_plan_kinds(planner):
    pl, dtype = planner
    seen = set()
    for n in list(range(201, 1001, 37)) + [255, 256, 257, 511, 512, 617, 719, 991, 997, 1000, 1024, 1234, 2047, 2048]:
        f = check_fft_algorithm(pl, n, DIRS[n % 2], dtype)
        seen.add(f.describe().split("{")[0])
    assert {"Direct", "Bluestein", "Rader", "Smooth"} <= seen


@pytest.mark.parametrize("n,desc", [(6, "Smooth{6=3x2}"), (1000, "Smooth{1000=5x5x5x8}"), (343, "Smooth{343=7x7x7}"),
                                    (1536, "Smooth{1536=3x16x16x2}"), (7, "Smooth{7=7}"), (105, "Smooth{105=7x5x3}"),
                                    (143, "Smooth{143=13x11}"), (961, "Smooth{961=31x31}"), (31, "Smooth{31=31}"),
                                    (1196, "Smooth{1196=23x13x4}"), (1131, "Smooth{1131=29x13x3}"), (323, "Smooth{323=19x17}")])
def test_smooth_plans(planner, n, desc):
    """Lengths whose prime factors are <= 31 run natively (run-time radix list; odd primes 11..31 use the
    symmetric prime butterfly), cf. RadixN + the hard-coded butterflies of src/algorithm/butterflies.rs."""
    pl, dtype = planner
    f = check_fft_algorithm(pl, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=5)
    assert f.describe() == desc
    check_fft_algorithm(pl, n, DIRS[1], dtype, control_kind=oracle.PLANNER, chunks=70)  # > F transforms: several CTAs


@pytest.mark.parametrize("n,desc", [(5000, "SmoothFourStep{50x100}"), (4225, "SmoothFourStep{65x65}"),
                                    (10000, "SmoothFourStep{100x100}"), (44100, "SmoothFourStep{210x210}"),
                                    (29791, "SmoothFourStep{31x961}"), (17017, "SmoothFourStep{119x143}")])
def test_smooth_four_step_plans(planner, n, desc):
    """Composite lengths above the one-pass limit whose prime factors are <= 31: two passes over a run-time radix
    list (the reference: MixedRadix / GoodThomas trees over its butterflies, src/plan.rs:508-607,
    src/algorithm/mixed_radix.rs:128-158) instead of Bluestein's four."""
    pl, dtype = planner
    f = check_fft_algorithm(pl, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=3)
    # f32: 10000 and 44100 run through compiled composite tiles (SmoothTileGeo: 100 x 100, 196 x 225)
    compiled = {10000: "SmoothFourStep{100x100,compiled}", 44100: "SmoothFourStep{196x225,compiled}"}
    assert f.describe() == (compiled.get(n, desc) if dtype == np.complex64 else desc)
    check_fft_algorithm(pl, n, DIRS[1], dtype, control_kind=oracle.PLANNER, chunks=2)
    if dtype == np.complex64 and n in compiled:  # the run-time-radix passes of the same split stay reachable (other kernels)
        a, b = (int(v) for v in desc[desc.index("{") + 1:-1].split("x"))
        if (a, b) != tuple(int(v) for v in compiled[n][compiled[n].index("{") + 1:compiled[n].index(",")].split("x")):
            g = check_fft_algorithm(pl, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=2, recipe=rb.Recipe.mixed_radix(a, b))
            assert g.describe() == desc


@pytest.mark.parametrize("n,batch,desc", [(10000, 70, "SmoothFourStep{100x100,compiled}"), (48000, 9, "SmoothFourStep{128x375,compiled}"),
                                          (100000, 5, "SmoothFourStep{250x400,compiled}"), (1000000, 2, "SmoothFourStep{1000x1000,compiled}")])
def test_compiled_composite_two_pass_plans(lib, n, batch, desc):
    """f32 two-pass plans whose pass lengths have compiled composite tiles (radix-3/5/7 stages in the CTA engine): ragged tiles
    that straddle transforms, several chunks over several streams, both directions, all entry points."""
    pl = rb.FftPlanner(np.complex64, lib=lib)
    for d in DIRS:
        f = check_fft_algorithm(pl, n, d, np.complex64, control_kind=oracle.PLANNER, chunks=batch if d == DIRS[0] else 1)
        assert f.describe() == desc


def test_compiled_tile_lengths_in_both_roles(lib):
    """Every compiled composite tile length as the column pass (first factor) and as the row pass (second factor) of a two-pass plan."""
    pl = rb.FftPlanner(np.complex64, lib=lib)
    ls = [64, 100, 125, 128, 196, 200, 225, 250, 256, 375, 400, 500, 512, 625, 1000, 1024]
    for a, b in zip(ls[:-1], ls[1:]):
        f = check_fft_algorithm(pl, a * b, DIRS[(a + b) % 2], np.complex64, control_kind=oracle.PLANNER, chunks=2, recipe=rb.Recipe.mixed_radix(a, b))
        assert f.describe() == "SmoothFourStep{%dx%d,compiled}" % (a, b)


@pytest.mark.parametrize("rdtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_real_fft_wrappers(lib, rdtype):
    """r2c / c2r of even lengths on top of the complex plans (tests/real_fft_cases.py)."""
    import real_fft_cases

    real_fft_cases.check_real_fft(rb.RealFftPlanner(rdtype, lib=lib), rdtype)


def test_fft_2d(planner):
    """2-D plans (tests/fft2d_cases.py): the width-point plan over the rows + one strided column pass."""
    import fft2d_cases

    pl, dtype = planner
    fft2d_cases.check_fft2d(pl, dtype, fft2d_cases.SHAPES[:-1])


def test_random_smooth_composites(planner):
    """Seeded random products of primes <= 31 between the one-pass limit and 600 000: every radix list / split the
    planner can produce for SmoothFourStep, against the f64 truth."""
    import random

    pl, dtype = planner
    rnd = random.Random(20260923)
    lo = 4096 if dtype == np.complex64 else 2048
    for t in range(16):
        n = 1
        while n <= lo:
            n *= rnd.choice([2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31] if rnd.random() < 0.5 else [2, 2, 3, 5])
        if n > 600000:
            continue
        inv = bool(t % 2)
        f = pl.plan_fft(n, DIRS[1] if inv else DIRS[0])
        assert f.describe().startswith("SmoothFourStep{"), (n, f.describe())
        x = signal(2 * n, dtype, seed=n)
        y = x.copy()
        f.process(y)
        assert rel_l2(y, truth(x, n, inv)) <= 4 * (5.96e-8 if dtype == np.complex64 else 1.11e-16) * np.log2(n), (n, f.describe())


def test_smooth_four_step_chunks_and_large(lib):
    pl = rb.FftPlanner(np.complex64, lib=lib)
    n, batch = 100000, 45  # 32 MiB of intermediate = 41 transforms per chunk: two chunks, the second ragged
    f = pl.plan_fft_forward(n)
    assert f.describe() == "SmoothFourStep{250x400,compiled}"
    f = pl.plan_fft_with_recipe(rb.Recipe.mixed_radix(160, 625), DIRS[0])  # a split without compiled tiles: the run-time-radix passes
    assert f.describe() == "SmoothFourStep{160x625}" and f.launches(batch) == 4
    x = signal(n * batch, np.complex64, seed=3)
    y = x.copy()
    f.process(y)
    for b in (0, 40, 41, 44):
        assert rel_l2(y[b * n:(b + 1) * n], truth(x[b * n:(b + 1) * n], n, False)) < 4 * 5.96e-8 * np.log2(n)
    f = pl.plan_fft_inverse(1000000)
    assert f.describe() == "SmoothFourStep{1000x1000,compiled}"
    x = signal(1000000, np.complex64, seed=4)
    y = x.copy()
    f.process(y)
    assert rel_l2(y, truth(x, 1000000, True)) < 4 * 5.96e-8 * np.log2(1000000)


@pytest.mark.parametrize("n,desc32,desc64", [
    (2048, "Direct{2048}", "Direct{2048}"), (4096, "Direct{4096}", "Direct{4096}"),
    (8192, "Direct{8192}", "Direct{8192}"), (1 << 14, "Direct{16384}", "FourStep{128x128}"),
    (1 << 15, "FourStep{128x256}", "FourStep{128x256}"), (1 << 16, "FourStep{256x256}", "FourStep{256x256}"),
    (1 << 17, "FourStep{256x512}", "FourStep{256x512}")])
def test_power_of_two_plans(planner, n, desc32, desc64):
    pl, dtype = planner
    f = check_fft_algorithm(pl, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=2)
    want = desc32 if dtype == np.complex64 else desc64
    # two-pass plans run as the single-launch dataflow kernel: "FourStep{N1xN2,flow,ring=W}"
    # (f32 default: the fused warp-specialised kernel, "FourStep{N1xN2,fused,ring=W}")
    assert f.describe() == want or f.describe().startswith(want[:-1] + ",flow,ring=") or f.describe().startswith(want[:-1] + ",fused,ring=")
    check_fft_algorithm(pl, n, DIRS[1], dtype, control_kind=oracle.PLANNER, chunks=1)


def test_largest_four_step_f32(lib):
    pl = rb.FftPlanner(np.complex64, lib=lib)
    f = check_fft_algorithm(pl, 1 << 20, DIRS[0], np.complex64, control_kind=oracle.PLANNER, chunks=1)
    assert f.describe().startswith("FourStep{1024x1024")


@pytest.mark.parametrize("check", plan_kinds.ALL, ids=[c.__name__[6:] for c in plan_kinds.ALL])
def test_round2_plan_kinds(planner, check):
    """General Rader, MixedRadix{r0 x Rader}, Good-Thomas, Bluestein over smooth lengths, caller-owned recipes (tests/plan_kinds.py)."""
    pl, dtype = planner
    check(pl, dtype)


@pytest.mark.parametrize("n,desc", [
    (65537, "Rader{n=65537,g=3,inner=FourStep{256x256}}"),          # BASELINE config 4
    (4099, "Bluestein{n=4099,M=16384,inner=FourStep{128x128}}"),
    (10007, "Bluestein{n=10007,M=32768,inner=FourStep{128x256}}"),
    (216569, "Bluestein{n=216569,M=524288,inner=FourStep{512x1024}}"),  # a 32-bit-overflow prime of raders_algorithm.rs:311-322
])
def test_large_convolution_plans(planner, n, desc):
    pl, dtype = planner
    f = check_fft_algorithm(pl, n, DIRS[0], dtype, control_kind=oracle.PLANNER, chunks=2)
    assert f.describe() == desc
    check_fft_algorithm(pl, n, DIRS[1], dtype, control_kind=oracle.PLANNER, chunks=1)


def test_fused_four_step_batches(lib):
    """f32 two-pass plans run as ONE launch of the fused warp-specialised kernel (fused.h); the replay harness takes the
    tiles in ticket order and checks that every tile runs exactly once and never before its dependency."""
    pl = rb.FftPlanner(np.complex64, lib=lib)
    n, batch = 1 << 16, 70
    x = signal(n * batch, np.complex64, seed=5)
    f = pl.plan_fft_forward(n)
    assert "fused" in f.describe() and f.launches(batch) == 1
    y = x.copy()
    f.process(y)
    for b in (0, 31, 32, 63, 64, 69):
        assert rel_l2(y[b * n:(b + 1) * n], truth(x[b * n:(b + 1) * n], n, False)) < 4 * 5.96e-8 * 16


@pytest.mark.parametrize("env", [{"B200FFT_FUSED": "0"}, {"B200FFT_FUSED": "0", "B200FFT_TMA_TILES": "0"}, {"B200FFT_FLOW": "1"},
                                 {"B200FFT_FLOW": "1", "B200FFT_FLOW_W": "2"},
                                 {"B200FFT_FLOW": "1", "B200FFT_FLOW_LOOKAHEAD": "3000"}, {"B200FFT_FUSED": "0", "B200FFT_NARROW": "1"},
                                 {"B200FFT_FUSED": "0", "B200FFT_NARROW": "1", "B200FFT_TMA_TILES": "0"},
                                 {"B200FFT_FUSED_W": "2"}, {"B200FFT_FUSED_LOOKAHEAD": "40"}, {"B200FFT_FUSED_LOOKAHEAD": "5000"},
                                 {"B200FFT_FUSED_TILED": "1"}, {"B200FFT_FUSED_BDIRECT": "63"}],
                         ids=["chunked-tma-tiles", "chunked-ldg-tiles", "flow", "flow-ring2", "flow-deep-lookahead", "narrow-tma-tiles",
                              "narrow-ldg-tiles", "fused-ring2", "fused-short-lookahead", "fused-deep-lookahead", "fused-tile-major-ring", "fused-direct-pass-b-output"])
def test_two_pass_variants_in_a_fresh_process(env):
    """The library reads its switches once per process: the chunked launch pairs (B200FFT_FUSED=0; TMA tiles or LDG/STG passes), the
    fused kernel with other ring sizes, and the single-launch dataflow kernel (B200FFT_FLOW=1, several ring sizes) are replayed in processes of their own; the
    replay harness also checks that every dataflow tile runs exactly once and never before its dependency."""
    import os
    import subprocess
    import sys

    from util import ROOT

    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = ROOT + os.pathsep + os.path.join(ROOT, "tests")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_variant_check.py")], env=e, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "VARIANT-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_host_pipeline_ring_wraps(lib):
    """The host-slice path stages 1 MiB chunks here (tests/util.py): 11 chunks through a ring of 4 device
    buffers, in place and out of place, with a ragged last chunk."""
    pl = rb.FftPlanner(np.complex64, lib=lib)
    n, batch = 1024, 1350  # 8 KiB per transform -> 128 per chunk -> 11 chunks, last one ragged
    f = pl.plan_fft_forward(n)
    x = signal(n * batch, np.complex64, seed=9)
    a = x.copy()
    f.process(a)
    b = np.zeros_like(x)
    f.process_outofplace_with_scratch(x.copy(), b)
    assert np.array_equal(a, b)
    for t in (0, 127, 128, 1000, 1349):
        assert rel_l2(a[t * n:(t + 1) * n], truth(x[t * n:(t + 1) * n], n, False)) < 4 * 5.96e-8 * 10


def test_error_behaviour_and_cache(planner):
    pl, dtype = planner
    check_error_behaviour(pl, dtype)
    check_planner_cache(pl)


def test_unsupported_lengths_fail_loudly(lib):
    pl = rb.FftPlanner(np.complex64, lib=lib)
    with pytest.raises(rb.FftError, match="not planned by this build"):
        pl.plan_fft_forward(1 << 25)
    with pytest.raises(rb.FftError, match="not planned by this build"):
        pl.plan_fft_forward((1 << 23) + 1)


def test_four_step_beyond_2_20(lib):
    """2^21 and 2^22 through the 2048-point tiles (up to 2^24 = 4096 x 4096 on the GPU suite)."""
    pl = rb.FftPlanner(np.complex64, lib=lib)
    for n, desc in [(1 << 21, "FourStep{1024x2048}"), (1 << 22, "FourStep{2048x2048}")]:
        f = pl.plan_fft_forward(n)
        assert f.describe().startswith(desc[:-1])
        x = signal(n, np.complex64, seed=21)
        y = x.copy()
        f.process(y)
        assert rel_l2(y, truth(x, n, False)) <= 4 * 5.96e-8 * np.log2(n)


def test_linearity_roundtrip_parseval(planner):
    pl, dtype = planner
    n = 1234
    a, b = signal(n, dtype, 1), signal(n, dtype, 2)
    f, fi = pl.plan_fft_forward(n), pl.plan_fft_inverse(n)
    fa, fb, fab = a.copy(), b.copy(), (a + 2 * b).astype(dtype)
    f.process(fa), f.process(fb), f.process(fab)
    tol = 64 * (5.96e-8 if dtype == np.complex64 else 1.11e-16)
    assert rel_l2(fab, fa + 2 * fb) < tol
    back = fa.copy()
    fi.process(back)
    assert rel_l2(back / n, a) < tol  # unnormalised both ways (src/lib.rs:81-85)
    assert abs(np.sum(np.abs(fa.astype(np.complex128)) ** 2) / n / np.sum(np.abs(a.astype(np.complex128)) ** 2) - 1) < tol
