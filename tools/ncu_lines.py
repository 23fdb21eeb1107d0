#!/usr/bin/env python3
"""Stall samples per CUDA source line of an .ncu-rep (needs -lineinfo + --import-source on): python tools/ncu_lines.py <rep> [top]"""
import csv, io, subprocess, sys, os
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, out = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        si = hdr.index("# Samples")
        stalls = sorted(((int(r[i] or 0), hdr[i][6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:2]
        out.append((int(r[si] or 0), int(r[hdr.index("Instructions Executed")] or 0), cur, r[0], r[1].strip(), stalls))
tot = sum(o[0] for o in out)
print(f"total samples {tot}")
for s, ex, f, ln, src, st in sorted(out, reverse=True)[:top]:
    print(f"{s * 100.0 / tot:5.1f}%  ex {ex:>10}  {f}:{ln:<4} {src[:100]:100s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
