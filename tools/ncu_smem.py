#!/usr/bin/env python3
"""Shared-memory wavefronts per CUDA source line of an .ncu-rep, excess (bank conflicts) first:
   python tools/ncu_smem.py <rep> [kernel-index] [top]   -- no GPU needed"""
import csv, io, os, subprocess, sys
path = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
cur, hdr, out, named = "?", None, [], False
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if not named: print(r[1][:160]); named = True
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        w = int(r[hdr.index("L1 Wavefronts Shared")] or 0); x = int(r[hdr.index("L1 Wavefronts Shared Excessive")] or 0)
        idl = int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)
        if w: out.append((x, w, idl, cur, r[0], r[1].strip()))
tw = sum(o[1] for o in out); tx = sum(o[0] for o in out)
print(f"shared wavefronts {tw}, excessive {tx} ({100.0 * tx / max(tw, 1):.1f}%)")
for x, w, idl, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"excess {x:>10} of {w:>10} (ideal {idl:>10})  {f}:{ln:<4} {src[:110]}")
