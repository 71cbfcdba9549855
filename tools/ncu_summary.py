#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(k) for k, _ in KEYS if k in hdr}
    name_i = hdr.index("Kernel Name")
    print(f"# ncu summary of `{path}` (ncu --set full --clock-control none; per launch, cold-cache, serialised)\n")
    print("| kernel | " + " | ".join(lbl for k, lbl in KEYS if k in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in data:
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v, u = r[idx[k]], units[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
        print("| `" + r[name_i].split("(")[0][-48:] + "` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
