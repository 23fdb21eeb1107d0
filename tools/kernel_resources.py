#!/usr/bin/env python3
"""Registers / spill stack / static shared memory of every kernel in the built objects (cuobjdump, no GPU needed):
   python tools/kernel_resources.py > profiles/<round>_kernel_resources.txt"""
import glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for obj in sorted(glob.glob(os.path.join(ROOT, "ckfft_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1); continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = dem.replace("ckb::", "").replace("(int)", "").replace("(bool)", "")
            rows.append((os.path.basename(obj), dem[:150], int(m.group(1)), int(m.group(2)), int(m.group(3))))
            name = None
print("# cuobjdump --dump-resource-usage over ckfft_b200/build/*.o (sm_100a): registers per thread, spill stack bytes, static shared bytes")
print("# (dynamic shared memory is set per launch: Cfg::SMEM_BYTES / TileCfg::SMEM_BYTES / SmallCfg::SMEM_BYTES)")
print("object | kernel | regs | stack | static_smem")
for r in rows:
    print(" | ".join(str(x) for x in r))
print(f"# {len(rows)} kernels, {sum(1 for r in rows if r[3] > 0)} with a spill stack")
