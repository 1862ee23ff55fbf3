"""Pipeline timeline of the fused two-pass kernel (B200FFT_FUSED_TRACE=1 makes the first 16 CTAs stamp %globaltimer at every
pipeline event into the tail of the workspace).  Usage: B200FFT_FUSED_TRACE=1 python tools/fused_trace.py LOG2N [batch]"""
import os
import re
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustfft_b200 as rb

CTAS, WORDS = 16, 4096
NAMES = {1: "P ticket", 2: "P stage-free", 3: "P load-queued", 4: "S tile-seen", 5: "S store-queued", 6: "S released",
         7: "C meta", 8: "C landed", 9: "C done"}


def main():
    lg = int(sys.argv[1])
    n = 1 << lg
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 28) // n
    f = rb.FftPlanner(np.complex64).plan_fft_forward(n)
    W = int(re.search(r"ring=(\d+)", f.describe()).group(1))
    x = torch.view_as_complex(torch.rand(batch * n, 2, device="cuda") * 10).contiguous()
    y = torch.empty_like(x)
    ws = torch.zeros(f.workspace_bytes(batch), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        f.process_device(x, out=y, workspace=ws)
    torch.cuda.synchronize()
    ctl = ((32 + 2 * W) * 4 + 255) // 256 * 256
    off = ctl + W * n * 8
    raw = ws[off: off + CTAS * 4 * WORDS * 8].cpu().numpy().view(np.uint64).reshape(CTAS, 4, WORDS)
    t0 = min(int(raw[c, r, 0] >> 8) for c in range(CTAS) for r in range(4) if raw[c, r, WORDS - 1] > 0)
    print(f"# {f.describe()} batch={batch}")
    for c in (0, 7):
        ev = []
        for r in range(4):
            cnt = int(raw[c, r, WORDS - 1])
            for k in range(cnt):
                w = int(raw[c, r, k])
                ev.append(((w >> 8) - t0, r, (w >> 4) & 15, w & 15))
        ev.sort()
        print(f"## CTA {c}: {len(ev)} events; window 60..85 us")
        for t, r, tag, arg in ev:
            if 60000 <= t <= 85000:
                print(f"{t / 1000:9.2f} us  role{r}  {NAMES.get(tag, tag):16s} {arg}")
    # durations over all traced CTAs
    load, comp, store, idle, seen, pwait, pissue = [], [], [], [], [], [], []
    for c in range(CTAS):
        last = {}
        ev = []
        for r in range(4):
            for k in range(int(raw[c, r, WORDS - 1])):
                w = int(raw[c, r, k])
                ev.append(((w >> 8), (w >> 4) & 15, w & 15))
        ev.sort()
        for t, tag, arg in ev:
            if tag == 1:
                last["tk"] = t
            elif tag == 2:
                if "tk" in last:
                    pwait.append(t - last["tk"])
                last["sf"] = t
            if tag == 3:
                if "sf" in last:
                    pissue.append(t - last["sf"])
                if ("rel", arg) in last:
                    idle.append(t - last[("rel", arg)])
                last[("q", arg)] = t
            elif tag == 8 and ("q", arg) in last:
                load.append(t - last[("q", arg)])
                last[("land", arg)] = t
            elif tag == 9 and ("land", arg) in last:
                comp.append(t - last[("land", arg)])
                last[("done", arg)] = t
            elif tag == 4 and ("done", arg) in last:
                seen.append(t - last[("done", arg)])
            elif tag == 5:
                last[("sq", arg)] = t
            elif tag == 6 and ("sq", arg) in last:
                store.append(t - last[("sq", arg)])
                last[("rel", arg)] = t
    for name, v in (("load queued->landed (incl. consumer pickup)", load), ("compute landed->done", comp), ("done->storer saw it", seen),
                    ("store queued->stage released", store), ("stage released->next load queued", idle),
                    ("producer: ticket in hand->stage free", pwait), ("producer: stage free->load queued", pissue)):
        v = np.array(v, dtype=np.float64) / 1000
        if len(v):
            print(f"{name:45s} n={len(v):6d} median {np.median(v):6.2f} us  p10 {np.percentile(v, 10):6.2f}  p90 {np.percentile(v, 90):6.2f}")


if __name__ == "__main__":
    main()
