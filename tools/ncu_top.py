"""Top stalled SASS instructions of an ncu report:  python tools/ncu_top.py report.ncu-rep [N] [kernel-regex]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sel = sys.argv[3] if len(sys.argv) > 3 else None
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
seen = set()
for blk in ([b for b in blocks if sel in b["name"]] if sel else blocks[:1]):
    if blk["name"] in seen:
        continue
    seen.add(blk["name"])
    h = blk["hdr"]
    si = h.index("# Samples")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si] or 0) for r in blk["rows"])
    print(blk["name"][:100], "total samples", tot)
    agg = {}
    for r in blk["rows"]:
        for i in stall:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("  stall mix:", ", ".join(f"{k[6:]} {v * 100 // max(1, tot)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    order = sorted(range(len(blk["rows"])), key=lambda i: -int(blk["rows"][i][si] or 0))[:top]
    for i in sorted(order):
        r = blk["rows"][i]
        s = int(r[si] or 0)
        why = sorted(((int(r[j] or 0), h[j][6:]) for j in stall), reverse=True)[:2]
        print(f"  {i:5d} {s * 100.0 / max(1, tot):5.1f}%  {r[1].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
