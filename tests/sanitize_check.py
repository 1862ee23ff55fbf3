"""Helper run under compute-sanitizer by tools/gpu_sanitize.sh (memcheck / racecheck / synccheck): one small exec of every
kernel family, checked against the f64 truth.  Tiny batches -- the sanitizer slows kernels down by 10-100x."""
import os
import sys

import numpy as np

import rustfft_b200 as rb
from rustfft_b200 import Recipe as R
from util import rel_l2, signal, strict_bound, truth


def run(pl, dtype, n, batch, recipe=None):
    for direction in (rb.FftDirection.Forward, rb.FftDirection.Inverse):
        inv = direction == rb.FftDirection.Inverse
        f = pl.plan_fft(n, direction) if recipe is None else pl.plan_fft_with_recipe(recipe, direction)
        x = signal(n * batch, dtype, seed=n)
        y = x.copy()
        f.process(y)
        assert rel_l2(y, truth(x, n, inv)) <= strict_bound(n, dtype), (n, inv, f.describe())
        print("ok", np.dtype(dtype).name, n, f.describe(), flush=True)


def main():
    quick = os.environ.get("SANITIZE_QUICK") == "1"
    p32, p64 = rb.FftPlanner(np.complex64), rb.FftPlanner(np.complex128)
    # Direct, Direct{16384}, two-pass (fused / chunked / flow per environment), Smooth, SmoothFourStep, fused Rader / Bluestein,
    # SmoothConv (Rader, MixedRadix x Rader, Bluestein over a smooth length), large convolution plans, Good-Thomas
    for n, batch in [(64, 33), (1024, 9), (4096, 3), (1 << 14, 2), (1 << 15, 5), (1 << 16, 3)]:
        run(p32, np.complex64, n, batch)
    if not quick:
        for n, batch in [(1000, 9), (5000, 3), (257, 9), (719, 5), (617, 11), (1234, 7), (1283, 3), (7681, 2), (65537, 2), (4099, 2), (1 << 20, 1)]:
            run(p32, np.complex64, n, batch)
        run(p32, np.complex64, 1200, 3, R.good_thomas(25, 48))
        run(p32, np.complex64, 44100, 2, R.good_thomas(196, 225))
        for n, batch in [(1024, 5), (1 << 15, 3), (1234, 5), (1000, 5), (10000, 2), (97, 9)]:
            run(p64, np.complex128, n, batch)
        # kernels added after the first sanitizer session: compiled composite tiles, cluster plans (DSMEM exchange, cluster barriers;
        # full and half tiles; Rader 65537 / Bluestein inside one cluster pass), 3-CTA smooth instantiations, real-FFT pack / unpack
        for n, batch in [(10000, 5), (44100, 3), (16000, 3), (1000000, 1), (1009, 5)]:
            run(p32, np.complex64, n, batch)
        for lg in (14, 15, 16, 17):
            run(p32, np.complex64, 1 << lg, 3, R.cluster(1 << lg))
        for lg in (14, 15, 16):
            run(p32, np.complex64, 1 << lg, 3, R.cluster(1 << lg, half_tiles=True))
        run(p32, np.complex64, 65537, 2, R.rader(65537, 1, R.cluster(65536)))
        run(p32, np.complex64, 6007, 2, R.bluestein(6007, R.cluster(16384)))
        for rdt in (np.float32, np.float64):
            rp = rb.RealFftPlanner(rdt)
            for n in (1024, 1000, 4098):
                f = rp.plan_fft(n)
                x = (np.random.default_rng(n).random(n * 3) * 10).astype(rdt)
                X = np.zeros(3 * (n // 2 + 1), np.complex64 if rdt == np.float32 else np.complex128)
                f.forward(x, X)
                assert rel_l2(X, np.fft.rfft(x.astype(np.float64).reshape(3, n), axis=1).ravel()) <= strict_bound(n, X.dtype)
                back = np.zeros_like(x)
                f.inverse(X, back)
                assert rel_l2(back / n, x) <= 2 * strict_bound(n, X.dtype)
                print("ok real", np.dtype(rdt).name, n, flush=True)
    print("SANITIZE-OK")


if __name__ == "__main__":
    sys.exit(main())
