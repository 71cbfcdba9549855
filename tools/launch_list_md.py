#!/usr/bin/env python
"""ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv launch list -> markdown table.

    python tools/launch_list_md.py gpurun_out/x/launches_train.csv > profiles/rNN_ncu_train_step_launches.md
"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = OrderedDict()
for r in rows:
    kid, name, metric, unit, val = r[0], r[4], r[12], r[13], float(r[14].replace(",", ""))
    d = per.setdefault(kid, {"name": name.split("(")[0].replace("void ", "")[-56:], "ms": 0.0, "rd": 0.0, "wr": 0.0})
    if metric.startswith("gpu__time_duration"):
        d["ms"] = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    elif metric.startswith("dram__bytes_read"):
        d["rd"] = val * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1e-9)
    elif metric.startswith("dram__bytes_write"):
        d["wr"] = val * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit, 1e-9)
agg = OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1
    a["ms"] += d["ms"]; a["rd"] += d["rd"]; a["wr"] += d["wr"]
tot = sum(a["ms"] for a in agg.values())
print("| kernel | launches | ms | share | dram read GB | dram write GB |\n|---|---|---|---|---|---|")
for n, a in agg.items():
    print(f"| `{n}` | {a['n']} | {a['ms']:.3f} | {100 * a['ms'] / tot:.1f} % | {a['rd']:.2f} | {a['wr']:.2f} |")
print(f"| **total** | {sum(a['n'] for a in agg.values())} | {tot:.3f} | 100 % | {sum(a['rd'] for a in agg.values()):.2f} | "
      f"{sum(a['wr'] for a in agg.values()):.2f} |")
