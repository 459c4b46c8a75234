"""Instruction mix of an ncu report (source page): python tools/ncu_inst.py report.ncu-rep [kernel-substring]
Per kernel: executed warp instructions by opcode, and the hottest SASS lines."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
sel = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks:
    if sel not in b["name"]:
        continue
    h = b["hdr"]
    ie, si = h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie] or 0) for r in b["rows"])
    ops = Counter()
    for r in b["rows"]:
        t = r[1].strip().split()
        op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
        ops[op.split(".")[0]] += int(r[ie] or 0)
    print(b["name"][:90], "warp instructions", tot)
    print("  ", ", ".join(f"{k} {v * 100 / tot:.1f}%" for k, v in ops.most_common(22)))
