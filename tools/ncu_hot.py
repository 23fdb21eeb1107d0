#!/usr/bin/env python3
"""Top stall sites of an .ncu-rep (SASS view): python tools/ncu_hot.py <rep> [top] -- no GPU needed."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first line: kernel name; second: header
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; data = [r for r in rows[1:] if len(r) == len(hdr)]
si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print(f"{lines[0][:200]}\ntotal samples {tot}, {len(data)} SASS instructions, executed {sum(int(r[ex] or 0) for r in data)}")
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda k: -int(data[k][si] or 0))[:top]
for k in sorted(order):
    r = data[k]
    why = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[si]) * 100.0 / tot:5.1f}%  ex {r[ex]:>9}  {r[src].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
