"""Helper run in a subprocess by tests/test_emu_parity.py: two-pass plans through the CPU replay library (tests/emu,
test infrastructure) under whatever B200FFT_* switches the parent put in the environment."""
import os
import sys

import numpy as np

import rustfft_b200 as rb
from util import emu_library, rel_l2, signal, strict_bound, truth


def main():
    pl = rb.FftPlanner(np.complex64, lib=emu_library())
    flow = os.environ.get("B200FFT_FLOW") == "1"
    fused = not flow and os.environ.get("B200FFT_FUSED") != "0"
    for n, batch in [(1 << 15, 70), (1 << 16, 37), (1 << 17, 9)]:
        for direction in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
            inv = direction == rb.FftDirection.Inverse
            f = pl.plan_fft(n, direction)
            assert ("flow" in f.describe()) == flow and ("fused" in f.describe()) == fused, f.describe()
            if flow or fused:
                assert f.launches(batch) == 1
            x = signal(n * batch, np.complex64, seed=n)
            y = x.copy()
            f.process(y)
            assert rel_l2(y, truth(x, n, inv)) <= strict_bound(n, np.complex64), (n, inv, f.describe())
    if os.environ.get("B200FFT_NARROW") == "1":
        # the 4- / 8-column tiles only exist for 512- and 1024-point passes: 2^19 = 512 x 1024, 2^20 = 1024 x 1024
        for n in (1 << 19, 1 << 20):
            for direction in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
                inv = direction == rb.FftDirection.Inverse
                f = pl.plan_fft(n, direction)
                x = signal(n * 3, np.complex64, seed=n)
                y = x.copy()
                f.process(y)
                assert rel_l2(y, truth(x, n, inv)) <= strict_bound(n, np.complex64), (n, inv, f.describe())
    if not flow and not fused:
        # batch larger than one L2 chunk: B200FFT_CHUNK_MB=32 (tests/util.py) -> 64 transforms of L2 budget at 2^16, split over the
        # four overlapped streams: 16 per chunk, four workspaces, ceil(70/16) = 5 chunks x 2 passes
        f = pl.plan_fft_forward(1 << 16)
        assert f.launches(70) == 10 and f.workspace_bytes(70) == 4 * 16 * (1 << 16) * 8
    if flow or fused:
        # fewer transforms than the look-ahead / than the ring, and a batch that wraps the ring twice
        n = 1 << 16
        f = pl.plan_fft_forward(n)
        for batch in (1, 2, 3, 33, 70):
            x = signal(n * batch, np.complex64, seed=batch)
            y = x.copy()
            f.process(y)
            assert rel_l2(y, truth(x, n, False)) <= strict_bound(n, np.complex64), batch
    print("VARIANT-OK")


if __name__ == "__main__":
    sys.exit(main())
