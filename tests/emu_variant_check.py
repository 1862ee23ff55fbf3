"""Helper run in a subprocess by tests/test_emu_parity.py: two-pass plans through the CPU replay library (tests/emu,
test infrastructure) under whatever B200FFT_* switches the parent put in the environment."""
import os
import sys

import numpy as np

import rustfft_b200 as rb
from util import emu_library, rel_l2, signal, strict_bound, truth


def main():
    pl = rb.FftPlanner(np.complex64, lib=emu_library())
    flow = os.environ.get("B200FFT_FLOW", "1") != "0"
    for n, batch in [(1 << 15, 70), (1 << 16, 37), (1 << 17, 9)]:
        for direction in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
            inv = direction == rb.FftDirection.Inverse
            f = pl.plan_fft(n, direction)
            assert ("flow" in f.describe()) == flow, f.describe()
            if not flow and n == 1 << 16:
                # B200FFT_CHUNK_MB=32 (tests/util.py) -> 64 transforms of L2 budget over two overlapped streams:
                # 32 per chunk, two workspaces, ceil(37/32) = 2 chunks x 2 passes
                assert f.launches(batch) == 4 and f.workspace_bytes(batch) == 2 * 32 * n * 8
            x = signal(n * batch, np.complex64, seed=n)
            y = x.copy()
            f.process(y)
            assert rel_l2(y, truth(x, n, inv)) <= strict_bound(n, np.complex64), (n, inv, f.describe())
    print("VARIANT-OK")


if __name__ == "__main__":
    sys.exit(main())
