"""Helper run in a subprocess by tests/test_gpu_parity.py: the library reads its tuning switches from the
environment once per process, so every alternative code path is checked in a process of its own."""
import sys

import numpy as np

import oracle
import rustfft_b200 as rb
from util import rel_l2, signal, strict_bound, truth


def main():
    pl = rb.FftPlanner(np.complex64)
    for n, batch in [(1024, 37), (4096, 9), (8192, 5), (1 << 15, 70), (1 << 17, 9), (1 << 20, 2), (65537, 3), (5000, 3)]:
        for direction in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
            inv = direction == rb.FftDirection.Inverse
            f = pl.plan_fft(n, direction)
            x = signal(n * batch, np.complex64, seed=n)
            y = x.copy()
            f.process(y)
            ref = truth(x, n, inv)
            want = oracle.fft(x[:n], n, inv)
            assert rel_l2(y, ref) <= strict_bound(n, np.complex64), (n, inv, f.describe())
            assert rel_l2(y[:n], want) <= 2 * strict_bound(n, np.complex64), (n, inv, f.describe())
    print("VARIANT-OK")


if __name__ == "__main__":
    sys.exit(main())
