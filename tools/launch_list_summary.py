#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import gzip
import re
import sys


def main():
    path = sys.argv[1]
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        if "run_kernel" in name:
            m = re.search(r"Geo<(\w+), (\d+), (\d+), (\d+)", name)
            kind = ("conv pass A" if "LoadColsConv" in name else "conv pass B" if "StoreTransposedConv" in name else
                    "pass A (cols)" if "LoadCols" in name else "pass B (rows->transposed)" if "StoreTransposed" in name else
                    "Bluestein fused" if "BluesteinKernel" in name else "Rader fused" if "RaderKernel" in name else "Direct")
            key = f"b2::run_kernel {kind} {m.group(1)} L={m.group(2)} E={m.group(3)} F={m.group(4)} grid={r[col['Grid Size']]} block={r[col['Block Size']]}"
        else:
            key = "(not ours) " + name[:70]
        v = float(r[col["Metric Value"]].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = sum(a[1] for k, a in agg.items() if not k.startswith("(not ours)"))
    print(f"# per-kernel aggregation of `{path}` (gpu__time_duration.sum; ncu serialises launches and flushes caches, compare SHARES)\n")
    print("| kernel | launches | total us | mean us | share of our kernels |\n|---|---|---|---|---|")
    for k, a in agg.items():
        share = f"{a[1] / ours * 100:.1f} %" if not k.startswith("(not ours)") else "-"
        print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {share} |")


if __name__ == "__main__":
    main()
