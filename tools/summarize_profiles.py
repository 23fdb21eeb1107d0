#!/usr/bin/env python3
"""Turn gpurun_out/launches_<tag>.csv and prof_c2c1024_<tag>.ncu-rep into the tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag> <round-name>"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list ----
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))) if len(r) > 10]
hdr, data = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in data:
    a = agg.setdefault((r[ki], r[gi]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list of `python bench.py --no-cpu-baseline` ({rnd}), B200, --clock-control none",
         f"# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_{tag}.csv python bench.py --no-cpu-baseline",
         "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
         "# the bench launches the kernel 320 times on the device-resident batch (grid = persistent CTAs) and 4 x 256 times on",
         "# 32 MiB chunks for the host-buffer (e2e) leg",
         f"# total launches {len(data)}, total kernel time {tot / 1e6:.1f} ms",
         "kernel | grid | launches | total_ms | mean_us | share"]
for (k, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k[:160]} | {g} | {c} | {t / 1e6:.3f} | {t / c / 1e3:.2f} | {t / tot:.4f}")
open(os.path.join(out_dir, f"{rnd}_launches_bench.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:10]))

# ---- full capture ----
rep = os.path.join(ROOT, "gpurun_out", f"prof_c2c1024_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'gpc__cycles_elapsed.avg.per_second']
kn = hdr.index("Kernel Name")
out = [f"# ncu --set full --clock-control none, top kernel of bench.py ({rnd})",
       f"# kernel: {data[0][kn]}",
       f"# command: ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 2 -o gpurun_out/prof_c2c1024_{tag} python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline",
       "# workload: 2^20 transforms of N=1024 per launch; algorithmic bytes per launch = 16*1024*2^20 = 17,179,869,184",
       "metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data)))]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        out.append(f"{k} | {units[i]} | " + " | ".join(r[i] for r in data))
ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1.0, "Kbyte": 1e3}[units[ir]]
traffic = [(float(r[ir]) + float(r[iw])) * scale for r in data]
mean = sum(traffic) / len(traffic)
out.append(f"derived: dram traffic per launch = {mean:.4e} B = {mean / 17179869184:.4f} x algorithmic bytes (no re-reads)")
out.append(f"derived: warp instructions per transform = {float(data[0][hdr.index('smsp__inst_executed.sum')]) / 2 ** 20:.0f}")
open(os.path.join(out_dir, f"{rnd}_c2c1024_ncu_full.txt"), "w").write("\n".join(out) + "\n")
json.dump({"c2c1024": round(mean), "_source": f"profiles/{rnd}_c2c1024_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full)"},
          open(os.path.join(out_dir, "roofline_traffic.json"), "w"), indent=1)
print("\n".join(out))
