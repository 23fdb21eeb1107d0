#!/usr/bin/env python3
"""Executed warp instructions by SASS opcode of an .ncu-rep: python tools/ncu_opmix.py <rep> [kernel-index] -- no GPU needed"""
import collections, csv, io, subprocess, sys
path = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; data = [r for r in rows[1:] if len(r) == len(hdr) and r != hdr]
ex = hdr.index("Instructions Executed"); src = hdr.index("Source")
mix = collections.Counter()
for r in data:
    toks = r[src].replace("@!", "@").split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    mix[op.split(".")[0]] += int(r[ex] or 0)
tot = sum(mix.values())
print(lines[0][:150]); print("total executed warp instructions", tot)
for op, n in mix.most_common(24):
    print(f"{op:10s} {n:>12d} {100.0 * n / tot:5.1f}%")
