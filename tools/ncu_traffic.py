"""Turns an ncu CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) of a short bench run
into profiles/r02_ncu_traffic.json: per kernel, launches / time / DRAM bytes of ONE step (the launches between the
last two pairs of cloud_box_kernel markers).

    python tools/ncu_traffic.py gpurun_out/r02_traffic.csv profiles/r02_ncu_traffic.json
"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
name_i, metric_i, unit_i, val_i = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= val_i:
        continue
    d = launch.setdefault(r[0], {"name": r[name_i]})
    v = float(r[val_i].replace(",", ""))
    u = r[unit_i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    d[r[metric_i]] = v * scale
seq = list(launch.values())
marks = [i for i, d in enumerate(seq) if "cloud_box_kernel" in d["name"]]
assert len(marks) >= 4, "need at least two steps in the capture"
step = seq[marks[-4]:marks[-2]]          # one full step: two encoders = two index builds
agg = {}
for d in step:
    base = re.sub(r"^void ", "", d["name"]).split("(")[0]
    base = re.sub(r"<.*", "", base).replace("sg4d::", "")
    a = agg.setdefault(base, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["ms"] += d.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
out = {"source": sys.argv[1], "note": "one step of `python bench.py` (8 scenes x 78 clouds x 80000 points) under ncu; "
       "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum; times are ncu's serialised cold-cache durations",
       "kernels": dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, a in out["kernels"].items():
    print(f"{k:36s} launches {a['launches']:4d}  {a['ms']:8.3f} ms  {a['dram_bytes'] / 1e9:8.2f} GB  "
          f"{a['dram_bytes'] / max(a['ms'], 1e-9) / 1e6:8.1f} GB/s")
