#!/usr/bin/env python3
"""Where a kernel's shared-memory loads sit relative to the arithmetic (is a gather issued up front or just in time?).
    cuobjdump -sass -fun <mangled> obj.o | python tools/sass_sched.py
Prints, per run of LDS separated by < 4 other instructions, its position and length."""
import re
import sys

ops = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops.append(m.group(2).split(".")[0])
runs, cur, gap = [], None, 0
for i, op in enumerate(ops):
    if op == "LDS":
        if cur is None:
            cur = [i, 1]
        else:
            cur[1] += 1
        gap = 0
    elif cur is not None:
        gap += 1
        if gap >= 4:
            runs.append(tuple(cur)); cur = None
    if op in ("BAR", "UBLKCP", "STG"):
        if not runs or runs[-1] != op:
            if cur is not None:
                runs.append(tuple(cur)); cur = None
            if not runs or runs[-1] != (i, op):
                runs.append((i, op))
out, last = [], None
for r in runs:
    if isinstance(r[1], str):
        if last != r[1]:
            out.append(r[1])
        last = r[1]
    else:
        out.append(f"LDSx{r[1]}@{r[0]}"); last = None
print(f"{len(ops)} instructions: " + " ".join(out))
