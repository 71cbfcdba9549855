#!/usr/bin/env python
"""Top source lines of one kernel of an .ncu-rep by warp-stall samples (needs --import-source on, -lineinfo).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n] [launch_index]
"""
import csv
import io
import subprocess
import sys


def main():
    path, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the output holds one table per launch: split on the header rows
    tables, cur, name = [], None, None
    for r in rows:
        if r and r[0] == "Function Name":
            name = r[1]
        if r and r[0] == "Line No":
            cur = {"hdr": r, "rows": [], "name": name}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    if not tables:
        print("no source table found")
        return
    t = tables[min(which, len(tables) - 1)]
    hdr = t["hdr"]
    i_samp = hdr.index("# Samples")
    i_inst = hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    src = [r for r in t["rows"] if r[2] == "-"]
    total = sum(int(r[i_samp] or 0) for r in src) or 1
    print(f"{t['name'][:100]}  tables={len(tables)} total samples={total}")
    src.sort(key=lambda r: -int(r[i_samp] or 0))
    for r in src[:top]:
        s = int(r[i_samp] or 0)
        stalls = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:3]
        st = " ".join(f"{h[6:]}={v}" for v, h in stalls if v)
        print(f"{100.0 * s / total:5.1f}%  inst={int(r[i_inst] or 0):>10}  L{r[0]:>4}  {r[1].strip()[:110]}   [{st}]")


if __name__ == "__main__":
    main()
