#!/usr/bin/env python
"""Sum DRAM traffic and durations of every kernel in an .ncu-rep (one forward captured by tools/profile_step.py)
and write profiles/traffic.json, which bench.py reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/prof_fwd.ncu-rep c4_n1000 profiles/traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

path, workload, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    i = hdr.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return [float(r[i].replace(",", "")) * scale.get(units[i], 1.0) for r in data]


rd, wr, ms = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
names = [r[hdr.index("Kernel Name")].split("(")[0][-40:] for r in data]
entry = {"source": os.path.basename(path), "launches": len(data), "dram_bytes": sum(rd) + sum(wr),
         "dram_read_bytes": sum(rd), "dram_write_bytes": sum(wr), "ncu_ms_total": sum(ms),
         "kernels": [{"kernel": n, "dram_bytes": a + b, "ncu_ms": t} for n, a, b, t in zip(names, rd, wr, ms)]}
db = json.load(open(out)) if os.path.exists(out) else {}
db[workload] = entry
json.dump(db, open(out, "w"), indent=1)
print(json.dumps({k: entry[k] for k in ("launches", "dram_bytes", "ncu_ms_total")}))
