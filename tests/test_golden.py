"""Golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py): the reference's own
known-answer tests plus oracle outputs on seeded inputs.  CPU: the oracle still reproduces them; GPU: the CUDA
path matches them."""
import os

import numpy as np
import pytest

import oracle
import rustfft_b200 as rb
from util import ROOT, mean_abs_err, rel_l2, signal, strict_bound

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
LENS = sorted({int(k.split("_")[-1]) for k in G.files if k.startswith("f32_fwd_")})
DT = {"f32": np.complex64, "f64": np.complex128}


def _input(name, n):
    return signal(n, DT[name], seed=1000 + n)


def test_oracle_reproduces_golden_bit_for_bit():
    for name in ("f32", "f64"):
        for n in LENS:
            x = _input(name, n)
            assert np.array_equal(oracle.fft(x, n, False), G[f"{name}_fwd_{n}"]), (name, n)
            assert np.array_equal(oracle.fft(x, n, True), G[f"{name}_inv_{n}"]), (name, n)


def test_reference_kats_are_in_the_golden_file():
    for i in range(4):
        sig, spec = G[f"kat_in_{i}"], G[f"kat_out_{i}"]
        assert mean_abs_err(oracle.fft(sig, len(sig), False, kind=oracle.DFT), spec) < 0.1  # src/algorithm/dft.rs:283-398


@pytest.mark.gpu
@pytest.mark.parametrize("name,dtype", [("f32", np.complex64), ("f64", np.complex128)])
def test_cuda_path_matches_golden(name, dtype):
    pl = rb.FftPlanner(dtype)
    for i in range(4):
        if dtype == np.complex64:
            sig = G[f"kat_in_{i}"].copy()
            pl.plan_fft_forward(len(sig)).process(sig)
            assert mean_abs_err(sig, G[f"kat_out_{i}"]) < 0.1
    for n in LENS:
        for key, d in (("fwd", rb.FftDirection.Forward), ("inv", rb.FftDirection.Inverse)):
            y = _input(name, n)
            pl.plan_fft(n, d).process(y)
            want = G[f"{name}_{key}_{n}"]
            assert mean_abs_err(y, want) < 0.1
            assert rel_l2(y, want) <= 2 * strict_bound(n, dtype), (name, n, key)


G2 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))


def _lens2(name):
    return sorted({int(k.split("_")[-1]) for k in G2.files if k.startswith(f"{name}_fwd_")})


def test_oracle_reproduces_golden_v2_bit_for_bit():
    for name in ("f32", "f64"):
        for n in _lens2(name):
            x = _input(name, n)
            assert np.array_equal(oracle.fft(x, n, False), G2[f"{name}_fwd_{n}"]), (name, n)
            assert np.array_equal(oracle.fft(x, n, True), G2[f"{name}_inv_{n}"]), (name, n)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dtype", [("f32", np.complex64), ("f64", np.complex128)])
def test_cuda_two_pass_plans_match_golden_v2(name, dtype):
    """Lengths of the two-pass plans (SmoothFourStep 5000 / 10000, TMA-tiled FourStep 2^15) against the committed
    oracle outputs."""
    pl = rb.FftPlanner(dtype)
    for n in _lens2(name):
        for key, d in (("fwd", rb.FftDirection.Forward), ("inv", rb.FftDirection.Inverse)):
            y = _input(name, n)
            pl.plan_fft(n, d).process(y)
            want = G2[f"{name}_{key}_{n}"]
            assert mean_abs_err(y, want) < 0.1
            assert rel_l2(y, want) <= 2 * strict_bound(n, dtype), (name, n, key)
