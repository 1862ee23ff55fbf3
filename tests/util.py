"""Shared helpers of the test-suite (signal generator, error metrics, library loaders)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EPS = {np.dtype(np.complex64): 5.96e-8, np.dtype(np.complex128): 1.11e-16}


def signal(n, dtype, seed=0):
    """The reference's test distribution: re, im ~ U[0, 10) (tests/accuracy.rs:84-95,
    src/test_utils.rs:23-34).  The exact rand-0.8 StdRng stream is not reproducible without that
    crate; pass/fail never depends on the particular values."""
    rng = np.random.default_rng(seed)
    return ((rng.random(n) + 1j * rng.random(n)) * 10).astype(dtype)


def mean_abs_err(a, b):
    """compare_vectors of the reference: mean |a-b| (must be < 0.1) -- src/test_utils.rs:36-43."""
    return float(np.mean(np.abs(a.astype(np.complex128) - b.astype(np.complex128))))


def rel_l2(a, b):
    a = a.astype(np.complex128)
    b = b.astype(np.complex128)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb else float(np.linalg.norm(a - b))


def truth(x, n, inverse):
    """f64 ground truth (numpy pocketfft), unnormalised like the reference (src/lib.rs:81-85)."""
    x = x.astype(np.complex128).reshape(-1, n)
    return (np.fft.ifft(x, axis=1) * n if inverse else np.fft.fft(x, axis=1)).ravel()


def strict_bound(n, dtype, factor=4.0):
    """relative L2 error <= 4 * eps_T * log2(N) against the f64 truth."""
    return factor * EPS[np.dtype(dtype)] * max(1.0, float(np.log2(max(n, 2))))


def emu_library():
    """The test-only CPU replay of the kernels behind the same C ABI (tests/emu)."""
    import __graft_entry__ as ge
    import rustfft_b200 as rb

    os.environ.setdefault("B200FFT_CHUNK_MB", "32")  # read once by the library; pins the chunking the tests assert
    os.environ.setdefault("B200FFT_HOST_CHUNK_MB", "1")  # small staging chunks: the host ring wraps in cheap tests
    return rb.Library(ge.build_emu())
