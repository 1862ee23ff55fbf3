"""A/B timing of alternative PLANS for the same length (caller-owned recipes, include/b200fft.h): each new plan kind of round 2
next to the plan it replaces.  Device-resident, ~1 GiB of signal per case (inputs larger than L2), CUDA events around `reps`
execs after 2 warm-ups; prints one row per (length, precision, recipe) with the fraction of the measured HBM roofline
(algorithmic bytes = one read + one write of the signal)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfft_b200 as rb
from rustfft_b200 import Recipe as R

HBM = 6487.4
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

AUTO = None
CASES = [
    # (length, [(label, recipe or AUTO)])
    (37, [("auto", AUTO), ("rader", R.rader(37)), ("bluestein-pow2", R.bluestein(37))]),
    (97, [("auto", AUTO), ("rader", R.rader(97)), ("bluestein-pow2", R.bluestein(97))]),
    (617, [("auto", AUTO), ("rader", R.rader(617)), ("bluestein-pow2", R.bluestein(617))]),
    (1009, [("auto", AUTO), ("rader", R.rader(1009)), ("bluestein-pow2", R.bluestein(1009))]),
    (2053, [("auto", AUTO), ("rader", R.rader(2053)), ("bluestein-pow2", R.bluestein(2053))]),
    (1234, [("auto", AUTO), ("mixedradix-2xrader617", R.rader(1234, 2)), ("bluestein-pow2-4096", R.bluestein(1234)),
            ("bluestein-smooth-2500", R.bluestein(1234, R.smooth(2500)))]),
    (1283, [("auto", AUTO), ("bluestein-pow2-4096", R.bluestein(1283)), ("bluestein-smooth-2592", R.bluestein(1283, R.smooth(2592)))]),
    (719, [("auto", AUTO), ("bluestein-pow2-2048", R.bluestein(719)), ("bluestein-smooth-1440", R.bluestein(719, R.smooth(1440)))]),
    (7681, [("auto", AUTO), ("rader-smooth4step", R.rader(7681)), ("bluestein-pow2", R.bluestein(7681))]),
    (112501, [("auto", AUTO), ("rader-smooth4step", R.rader(112501)), ("bluestein-pow2", R.bluestein(112501))]),
    (65537, [("auto", AUTO), ("bluestein-pow2", R.bluestein(65537)), ("rader-cluster-8", R.rader(65537, 1, R.cluster(65536)))]),
    (20011, [("auto", AUTO), ("bluestein-cluster-8", R.bluestein(20011, R.cluster(65536)))]),
    (6007, [("auto", AUTO), ("bluestein-cluster-2", R.bluestein(6007, R.cluster(16384)))]),
    (4099, [("auto", AUTO), ("bluestein-pow2-16384", R.bluestein(4099, R.pow2(16384))), ("bluestein-smooth-8232", R.bluestein(4099, R.mixed_radix(84, 98)))]),
    (10000, [("auto", AUTO), ("smooth4step-100x100", R.mixed_radix(100, 100)), ("goodthomas-16x625", R.good_thomas(16, 625))]),
    (44100, [("auto", AUTO), ("smooth4step-210x210", R.mixed_radix(210, 210)), ("goodthomas-196x225", R.good_thomas(196, 225))]),
    (48000, [("auto", AUTO), ("goodthomas-128x375", R.good_thomas(128, 375))]),
    (1200, [("auto", AUTO), ("goodthomas-25x48", R.good_thomas(25, 48))]),
    (1000000, [("auto", AUTO)]),
    (100000, [("auto", AUTO)]),
    (16000, [("auto", AUTO)]), (62500, [("auto", AUTO)]), (200000, [("auto", AUTO)]), (390625, [("auto", AUTO)]),
    (1000, [("auto", AUTO)]), (3600, [("auto", AUTO)]), (2187, [("auto", AUTO)]), (961, [("auto", AUTO)]),
    # single pass on a thread-block cluster (DSMEM transpose) next to the default plan of the same length (f32)
    (1 << 14, [("auto", AUTO), ("cluster-2", R.cluster(1 << 14)), ("cluster-4-half-tiles", R.cluster(1 << 14, True))]),
    (1 << 15, [("auto", AUTO), ("cluster-4", R.cluster(1 << 15)), ("cluster-8-half-tiles", R.cluster(1 << 15, True))]),
    (1 << 16, [("auto", AUTO), ("cluster-8", R.cluster(1 << 16)), ("cluster-16-half-tiles", R.cluster(1 << 16, True))]),
    (1 << 17, [("auto", AUTO), ("cluster-16", R.cluster(1 << 17))]),
]


def main():
    only = [int(a) for a in sys.argv[1:]]
    rows = []
    for dtype in (np.complex64, np.complex128):
        esz = 8 if dtype == np.complex64 else 16
        pl = rb.FftPlanner(dtype)
        tdt = torch.complex64 if dtype == np.complex64 else torch.complex128
        for n, alts in CASES:
            if only and n not in only:
                continue
            batch = max(1, (1 << (32 if n >= (1 << 14) and (n & (n - 1)) == 0 else 30)) // (n * esz))
            x = torch.empty(batch * n, dtype=tdt, device="cuda")
            torch.view_as_real(x).uniform_(0, 10)
            y = torch.empty_like(x)
            for label, rc in alts:
                try:
                    f = pl.plan_fft_forward(n) if rc is None else pl.plan_fft_with_recipe(rc, rb.FftDirection.Forward)
                except rb.FftError as e:
                    print(f"{np.dtype(dtype).name} n={n} {label}: not plannable ({e})", flush=True)
                    continue
                ws = torch.empty(max(16, f.workspace_bytes(batch)), dtype=torch.uint8, device="cuda")
                for _ in range(2):
                    f.process_device(x, out=y, workspace=ws)
                torch.cuda.synchronize()
                reps = 5
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    f.process_device(x, out=y, workspace=ws)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                frac = 2.0 * esz * n * batch / ms / 1e6 / HBM
                gflops = 5.0 * n * np.log2(n) * batch / ms / 1e6
                rows.append({"dtype": np.dtype(dtype).name, "n": n, "label": label, "plan": f.describe(), "batch": batch, "ms": round(ms, 4),
                             "gflops": round(gflops, 1), "frac": round(frac, 4)})
                print(f"{np.dtype(dtype).name:10s} n={n:8d} batch={batch:8d} {label:24s} {ms:9.4f} ms  frac={frac:.3f}  {f.describe()}", flush=True)
                del ws
            del x, y
            torch.cuda.empty_cache()
    out = os.environ.get("AB_PLANS_OUT")
    if out:
        json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
