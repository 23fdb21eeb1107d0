#!/usr/bin/env python3
"""Static opcode mix of the main loop (largest backward-branch span) of a kernel's SASS.
    cuobjdump -sass -fun <mangled> obj.o | python tools/sass_loop.py"""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
best = None
for a, op, rest in ins:
    if op.startswith("BRA"):
        t = re.search(r"0x([0-9a-f]+)", rest)
        if t and int(t.group(1), 16) < a:
            span = (a - int(t.group(1), 16), int(t.group(1), 16), a)
            if best is None or span > best:
                best = span
if best is None:
    print("no loop"); sys.exit()
_, lo, hi = best
body = [op for a, op, _ in ins if lo <= a <= hi]
mix = collections.Counter(op.split(".")[0] for op in body)
print(f"total {len(ins)}  loop {len(body)} instructions [{lo:#x}, {hi:#x}]")
print("  ".join(f"{k}:{v}" for k, v in mix.most_common(24)))
