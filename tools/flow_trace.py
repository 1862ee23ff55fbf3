"""Per-phase timeline of the dataflow four-step kernel (B2_FLOW_TRACE build, B200FFT_FLOW=1): thread 0 of the first 32
CTAs stamps %globaltimer at every phase boundary of every tile; this prints the mean time between consecutive stamps
for pass-A and pass-B tiles in steady state."""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfft_b200 as rb

CTAS, WORDS = 32, 4096
pl = rb.FftPlanner(np.complex64)
for lg, batch in [(20, 512), (16, 8192)]:
    n = 1 << lg
    f = pl.plan_fft_forward(n)
    x = torch.empty(n * batch, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(0, 10)
    y = torch.empty_like(x)
    nws = f.workspace_bytes(batch)
    ws = torch.zeros(nws, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        f.process_device(x, out=y, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f.process_device(x, out=y, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"== {f.describe()} batch={batch}: {ms:.3f} ms, frac={16.0 * n * batch / ms / 1e6 / 6487.4:.3f}")
    tr = ws[nws - CTAS * WORDS * 8:].cpu().numpy().view(np.uint64).reshape(CTAS, WORDS)
    seg = {1: collections.defaultdict(list), 2: collections.defaultdict(list)}
    period = {1: [], 2: []}
    waits = 0
    ntiles = 0
    for c in range(CTAS):
        cnt = int(tr[c, WORDS - 1])
        st = [(int(v) >> 8, int(v) & 0xFF) for v in tr[c, :cnt]]
        tiles, cur = [], None
        for t, tag in st:
            if tag in (1, 2, 3):
                if cur:
                    tiles.append(cur)
                cur = [(t, tag)]
            elif cur is not None:
                cur.append((t, tag))
        if cur:
            tiles.append(cur)
        for i, tl in enumerate(tiles[3:-2], start=3):
            kind = tl[0][1]
            if kind == 3:
                continue
            ntiles += 1
            for (t0, g0), (t1, g1) in zip(tl[:-1], tl[1:]):
                seg[kind][(g0, g1)].append(t1 - t0)
                if g1 == 5:
                    waits += 1
            nxt = tiles[i + 1][0][0]
            period[kind].append(nxt - tl[0][0])
        if c == 0:
            print("   CTA0 sample:", " ".join(f"{tag:02x}+{t - tiles[5][0][0]}" for t, tag in tiles[5] + tiles[6]))
    for kind, name in ((1, "A"), (2, "B")):
        print(f"  pass {name}: mean tile period {np.mean(period[kind]):.0f} ns over {len(period[kind])} tiles")
        for k, v in seg[kind].items():
            print(f"     {k[0]:02x}->{k[1]:02x}: mean {np.mean(v):7.0f} ns  p90 {np.percentile(v, 90):7.0f}  n={len(v)}")
    print(f"  tiles that waited for a dependency: {waits} of {ntiles}")
