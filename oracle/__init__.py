"""ctypes front-end of the CPU oracle (oracle/rustfft_scalar_oracle.cpp).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; rustfft_b200 never does.

The oracle is a C++ restatement of RustFFT's scalar path (the reference is Rust and cannot be
compiled here: no rustc/cargo in the image).  Parity pinned by the reference's own known-answer
tests, see tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

PLANNER, CONTROL, DFT, RADIX4 = 0, 1, 2, 3


def build(force: bool = False) -> None:
    """Compile both oracle libraries (parity: -ffp-contract=off; fast: -O3) with oracle/Makefile."""
    src = os.path.join(_HERE, "rustfft_scalar_oracle.cpp")
    libs = [os.path.join(_BUILD, n) for n in ("liboracle_parity.so", "liboracle_fast.so")]
    stale = force or any(
        (not os.path.exists(p)) or os.path.getmtime(p) < os.path.getmtime(src) for p in libs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "-j2"], check=True, capture_output=True)


_libs: dict = {}


def _lib(fast: bool = False):
    key = "fast" if fast else "parity"
    if key not in _libs:
        build()
        lib = ctypes.CDLL(os.path.join(_BUILD, f"liboracle_{key}.so"))
        u64, i32 = ctypes.c_uint64, ctypes.c_int
        lib.oracle_fft_f32.argtypes = [i32, u64, i32, ctypes.c_void_p, u64, i32]
        lib.oracle_fft_f64.argtypes = [i32, u64, i32, ctypes.c_void_p, u64, i32]
        lib.oracle_time_f32.argtypes = [i32, u64, i32, ctypes.c_void_p, u64, i32, i32]
        lib.oracle_time_f32.restype = ctypes.c_double
        lib.oracle_bench_f32.argtypes = [i32, u64, u64, i32, i32, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_describe_plan.argtypes = [u64, ctypes.c_char_p, u64]
        lib.oracle_modular_exponent.argtypes = [u64, u64, u64]
        lib.oracle_modular_exponent.restype = u64
        lib.oracle_primitive_root.argtypes = [u64]
        lib.oracle_primitive_root.restype = u64
        lib.oracle_distinct_prime_factors.argtypes = [u64, ctypes.c_void_p, i32]
        lib.oracle_prime_factors.argtypes = [u64, ctypes.c_void_p, i32]
        lib.oracle_partition_factors.argtypes = [u64, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_twiddle_f64.argtypes = [u64, u64, i32, ctypes.c_void_p]
        lib.oracle_twiddle_f32.argtypes = [u64, u64, i32, ctypes.c_void_p]
        lib.oracle_is_prime.argtypes = [u64]
        _libs[key] = lib
    return _libs[key]


def fft(x: np.ndarray, n: int, inverse: bool = False, kind: int = PLANNER, threads: int = 1,
        fast: bool = False) -> np.ndarray:
    """Transform every contiguous chunk of `n` elements of `x` (complex64 / complex128), like
    Fft::process on a buffer of len = batch*n.  Returns a new array of the same shape."""
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.complex64, np.complex128):
        raise TypeError("oracle.fft wants complex64 or complex128")
    out = x.copy()
    if n == 0 or out.size == 0:
        return out
    if out.size % n:
        raise ValueError("buffer length must be a multiple of the FFT length")
    batch = out.size // n
    fn = _lib(fast).oracle_fft_f32 if out.dtype == np.complex64 else _lib(fast).oracle_fft_f64
    rc = fn(kind, n, int(inverse), out.ctypes.data, batch, threads)
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")
    return out


def time_f32(n: int, batch: int, threads: int, reps: int = 1, kind: int = PLANNER, seed: int = 1) -> float:
    """Seconds for `reps` passes over a batch of f32 transforms on `threads` host threads (fast build).
    Plans are built outside the timed region, as benches/bench_rustfft.rs:43-54 does."""
    rng = np.random.default_rng(seed)
    buf = (rng.random(2 * n * batch, dtype=np.float32) * 10).astype(np.float32)
    t = _lib(True).oracle_time_f32(kind, n, 0, buf.ctypes.data, batch, threads, reps)
    if t < 0:
        raise RuntimeError("oracle_time_f32 failed")
    return t


def bench_f32(n: int, batch: int, threads: int, reps: int, buf: np.ndarray = None, kind: int = PLANNER) -> list:
    """Per-pass seconds of `reps` timed passes over `batch` f32 transforms of length n on `threads` pinned workers (fast build).
    Workers are created once, first-touch and fill their own slice of `buf` (complex64, >= batch*n elements; allocated
    uninitialised here when None), and every pass is timed between two barriers -- see oracle_bench_f32."""
    if buf is None:
        buf = np.empty(n * batch, dtype=np.complex64)
    assert buf.dtype == np.complex64 and buf.size >= n * batch
    times = (ctypes.c_double * reps)()
    rc = _lib(True).oracle_bench_f32(kind, n, batch, threads, reps, buf.ctypes.data, times)
    if rc != 0:
        raise RuntimeError("oracle_bench_f32 failed")
    return [float(times[i]) for i in range(reps)]


def describe_plan(n: int) -> str:
    buf = ctypes.create_string_buffer(1 << 16)
    rc = _lib().oracle_describe_plan(n, buf, len(buf))
    if rc < 0:
        raise RuntimeError("plan description too long")
    return buf.value.decode()


def modular_exponent(b: int, e: int, m: int) -> int:
    return int(_lib().oracle_modular_exponent(b, e, m))


def primitive_root(p: int) -> int:
    return int(_lib().oracle_primitive_root(p))


def distinct_prime_factors(n: int) -> list:
    out = (ctypes.c_uint64 * 64)()
    k = _lib().oracle_distinct_prime_factors(n, out, 64)
    return [int(out[i]) for i in range(k)]


def prime_factors(n: int) -> dict:
    out = (ctypes.c_uint64 * 128)()
    k = _lib().oracle_prime_factors(n, out, 128)
    v = [int(out[i]) for i in range(k)]
    other = [(v[5 + 2 * i], v[6 + 2 * i]) for i in range(v[4])]
    return {"p2": v[0], "p3": v[1], "total": v[2], "distinct": v[3], "other": other}


def partition_factors(n: int):
    l, r = ctypes.c_uint64(), ctypes.c_uint64()
    if _lib().oracle_partition_factors(n, ctypes.byref(l), ctypes.byref(r)) != 0:
        raise ValueError("prime or < 2")
    return int(l.value), int(r.value)


def twiddle(idx: int, n: int, inverse: bool = False, dtype=np.complex128):
    if dtype == np.complex128:
        o = (ctypes.c_double * 2)()
        _lib().oracle_twiddle_f64(idx, n, int(inverse), o)
    else:
        o = (ctypes.c_float * 2)()
        _lib().oracle_twiddle_f32(idx, n, int(inverse), o)
    return dtype(complex(o[0], o[1]))


def is_prime(n: int) -> bool:
    return bool(_lib().oracle_is_prime(n))
