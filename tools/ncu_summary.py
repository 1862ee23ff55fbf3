#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [algorithmic_bytes_per_launch ...] > profiles/xyz.md"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__inst_executed_pipe_fma.sum", "fma pipe inst"),
    ("smsp__inst_executed.sum", "inst executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occ %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{rep}`\n")
    for r in data:
        name = r[col["Kernel Name"]]
        short = re.sub(r"b2::", "", name)[:200]
        print(f"## launch {r[col['ID']]}: `{short}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in col:
                print(f"| {label} (`{k}`) | {r[col[k]]} | {units[col[k]]} |")
        # warp stall reasons, whatever this ncu version calls them: the eight largest "...issue_stalled_<reason>..." columns
        stalls = []
        for h, i in col.items():
            m = re.search(r"issue_stalled_([a-z_]+?)(?:_per_|\.|$)", h)
            if m and ("per_warp_active" in h or "ratio" in h):
                try:
                    stalls.append((float(r[i].replace(",", "")), m.group(1), h, units[i]))
                except ValueError:
                    pass
        seen = set()
        for v, reason, h, u in sorted(stalls, reverse=True):
            if reason in seen:
                continue
            seen.add(reason)
            if len(seen) > 8:
                break
            print(f"| stall {reason} (`{h}`) | {v:.3f} | {u} |")
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", ""))
            ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * mult.get(ur, 1) + wr * mult.get(uw, 1)
            dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            du = units[col["gpu__time_duration.sum"]]
            dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(du, 1e-9)
            print(f"| **traffic = dram read + write** | {tot:.4g} | byte |")
            print(f"| **dram GB/s under ncu (serialised, cold)** | {tot / dur_s / 1e9:.1f} | GB/s |")
        except Exception as e:  # noqa
            print(f"| traffic | n/a ({e}) | |")
        print()


if __name__ == "__main__":
    main()
