#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + instructions per source line (needs -lineinfo).
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [elements_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
elems = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
vals = rows[2 + which]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?", " (result", which, "of", len(rows) - 2, ")")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for k in keys:
    if k in d:
        print(f"{k:90s} {d[k][0]:>16s} {d[k][1]}")
if elems and "smsp__inst_executed.sum" in d:
    inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
    print(f"lane-instructions per element: {inst * 32 / elems:.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# the source page concatenates one block per profiled kernel: keep block `which`
blocks = src.split('"Kernel Name"')
if len(blocks) > 1:
    src = '"Kernel Name"' + blocks[1 + min(which, len(blocks) - 2)]
rows = list(csv.reader(io.StringIO(src)))
hdr = None
agg = collections.OrderedDict()
total = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or "Source" not in hdr:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        continue
    try:
        n = int(r[hdr.index("Instructions Executed")])
        st = int(r[hdr.index("Warp Stall Sampling (All Samples)")])
    except ValueError:
        continue
    line = r[0]
    if not line:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, [0, 0, r[hdr.index("Source")][:90]])
    a[0] += n
    a[1] += st
tot = sum(v[0] for v in agg.values())
print(f"\ninstructions attributed to source lines: {tot} (each SASS view may repeat; use shares)")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    pe = f"{v[0] * 32 / elems:6.1f}/elem" if elems else ""
    print(f"{f:16s} L{l:>4} inst={v[0]:>9} {v[0] / tot * 100:5.1f}% {pe} stall={v[1]:>6} | {v[2]}")
