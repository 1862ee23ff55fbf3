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
    dram = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        if len(r) < len(hdr):
            continue
        metric = r[col["Metric Name"]]
        if metric in dram:
            if "run_" in r[col["Kernel Name"]]:
                dram[metric] += float(r[col["Metric Value"]].replace(",", "")) * mult.get(r[col["Metric Unit"]], 1.0)
            continue
        if metric != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]]
        if "run_fused" in name:
            # fused single-launch four-step: the two tile kernels it runs (pass A L1-point columns, pass B L2-point rows)
            g = re.findall(r"Geo<(\w+), (\d+), (\d+), (\d+)", name)
            key = (f"b2::run_fused fused four-step (both passes, one persistent launch) {g[0][0]} {g[0][1]}x{g[1][1] if len(g) > 1 else g[0][1]} "
                   f"grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
        elif "run_cluster" in name:
            g = re.findall(r"Geo<(\w+), (\d+), (\d+), (\d+)", name)
            key = f"b2::run_cluster cluster plan {g[0][0]} L={g[0][1]} grid={r[col['Grid Size']]} block={r[col['Block Size']]}"
        elif "TmaTileKernel" in name or "run_flow" in name:
            m = re.search(r"Geo<(\w+), (\d+), (\d+), (\d+)", name)
            role = re.search(r">, \d, \d, (\d), \d>", name)
            kind = "dataflow four-step (both passes)" if "run_flow" in name else (
                "TMA pass A (cols)" if role and role.group(1) == "0" else "TMA pass B (rows->transposed)")
            key = f"b2::{'run_flow' if 'run_flow' in name else 'run_kernel_tma'} {kind} {m.group(1)} L={m.group(2)} E={m.group(3)} F={m.group(4)} grid={r[col['Grid Size']]} block={r[col['Block Size']]}"
        elif "run_kernel" in name or "run_pipelined" in name:
            m = re.search(r"Geo<(\w+), (\d+), (\d+), (\d+)", name)
            kind = ("conv pass A" if "LoadColsConv" in name else "conv pass B" if "StoreTransposedConv" in name else
                    "pass A (cols)" if "LoadCols" in name else "pass B (rows->transposed)" if "StoreTransposed" in name else
                    "Bluestein fused" if "BluesteinKernel" in name else "Rader fused" if "RaderKernel" in name else "Direct")
            key = f"b2::{'run_pipelined' if 'run_pipelined' in name else 'run_kernel'} {kind} {m.group(1)} L={m.group(2)} E={m.group(3)} F={m.group(4)} grid={r[col['Grid Size']]} block={r[col['Block Size']]}"
        else:
            key = "(not ours) " + name[:70]
        v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[col["Metric Unit"]], 1.0)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = sum(a[1] for k, a in agg.items() if not k.startswith("(not ours)"))
    print(f"# per-kernel aggregation of `{path}` (gpu__time_duration.sum; ncu serialises launches and flushes caches, compare SHARES)\n")
    print("| kernel | launches | total us | mean us | share of our kernels |\n|---|---|---|---|---|")
    for k, a in agg.items():
        share = f"{a[1] / ours * 100:.1f} %" if not k.startswith("(not ours)") else "-"
        print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {share} |")
    if dram["dram__bytes_read.sum"] or dram["dram__bytes_write.sum"]:
        tot = dram["dram__bytes_read.sum"] + dram["dram__bytes_write.sum"]
        print(f"\nDRAM traffic of our kernels over the profiled launches: read {dram['dram__bytes_read.sum']:.4g} B + "
              f"write {dram['dram__bytes_write.sum']:.4g} B = {tot:.4g} B")
        import json, os
        json.dump({"dram_bytes_per_step": tot, "dram_read": dram["dram__bytes_read.sum"], "dram_write": dram["dram__bytes_write.sum"],
                   "source": path}, open(os.path.join(os.path.dirname(path) or ".", "traffic.json"), "w"))


if __name__ == "__main__":
    main()
