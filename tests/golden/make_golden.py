#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz and golden_v2.npz.

The reference (RustFFT) cannot be executed in this environment (no Rust toolchain), so the golden vectors
are (a) the reference's own known-answer tests, transcribed (src/algorithm/dft.rs:283-398), and (b) outputs
of the CPU oracle -- the C++ restatement of the reference's scalar planner path, pinned by those KATs and by
an f64 numpy truth in tests/test_oracle.py -- on seeded inputs with the reference's test distribution
(U[0,10) re/im, tests/accuracy.rs:84-95), for the lengths the BASELINE configs and the reference's unit tests
exercise.  One transform per length, both directions, both precisions; inputs are not stored, they are
`tests/util.py::signal(n, dtype, seed=1000 + n)`.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from util import signal  # noqa: E402

LENS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 16, 17, 24, 27, 31, 32, 59, 64, 97, 100, 128, 257, 360, 617, 1000, 1024,
        1234]


LENS2 = [5000, 10000, 32768]  # SmoothFourStep{50x100}, SmoothFourStep{100x100}, FourStep{128x256}


def main():
    out = {}
    for name, dtype in (("f32", np.complex64), ("f64", np.complex128)):
        for n in LENS:
            x = signal(n, dtype, seed=1000 + n)
            out[f"{name}_fwd_{n}"] = oracle.fft(x, n, False)
            out[f"{name}_inv_{n}"] = oracle.fft(x, n, True)
    # the reference's hand-written KATs (signal, spectrum), f32
    kats = [
        ([1, -1], [0, 2]),
        ([1 + 1j, 2 - 3j, -1 + 4j], [2 + 2j, -5.562177 - 2.098076j, 6.562178 + 3.09807j]),
        ([1j, 2.5 - 3j, -1 - 1j, 4], [5.5 - 3j, -2 + 3.5j, -7.5 + 3j, 4 + 0.5j]),
        ([1 + 1j, 2 + 2j, 3 + 3j, 4 + 4j, 5 + 5j, 6 + 6j], [21 + 21j, -8.16 + 2.16j, -4.76 - 1.24j, -3 - 3j, -1.24 - 4.76j, 2.16 - 8.16j]),
    ]
    for i, (sig, spec) in enumerate(kats):
        out[f"kat_in_{i}"] = np.array(sig, dtype=np.complex64)
        out[f"kat_out_{i}"] = np.array(spec, dtype=np.complex64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")
    # second file: lengths of the two-pass plans (TMA-tiled FourStep, SmoothFourStep), added with those plans
    out2 = {}
    for name, dtype, lens in (("f32", np.complex64, LENS2), ("f64", np.complex128, LENS2[:1])):
        for n in lens:
            x = signal(n, dtype, seed=1000 + n)
            out2[f"{name}_fwd_{n}"] = oracle.fft(x, n, False)
            out2[f"{name}_inv_{n}"] = oracle.fft(x, n, True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz"), **out2)
    print("wrote", len(out2), "arrays")


if __name__ == "__main__":
    main()
