#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + SASS regions (contiguous instructions with similar
execution counts) with instruction mix, stall samples and shared-memory wavefronts.
usage: python tools/ncu_regions.py report.ncu-rep [min_share_percent]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    print("=" * 100)
    print(r[hdr.index("Kernel Name")])
    for k in keys:
        if k in hdr:
            print("  %-95s %16s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
kern, data, order = None, collections.defaultdict(list), []
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1][:60]
        if kern not in order:
            order.append(kern)
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if kern and len(r) > 10:
        data[kern].append(r)
iE, iS, iW, iWI, iT = (h.index(x) for x in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                             "Avg. Predicated-On Threads Executed"))
def opname(t):
    p = t.split()
    o = p[1] if p[0].startswith("@") else p[0]
    return o.split(".")[0]
for k in order:
    rs = data[k]
    # ncu repeats the listing when several launches of the same kernel are in the report
    addr0 = rs[0][0]
    reps = sum(1 for r in rs if r[0] == addr0)
    rs = rs[: len(rs) // reps]
    tot = sum(int(r[iE]) for r in rs)
    tots = sum(int(r[iS]) for r in rs)
    print("=" * 100)
    print(k, " total %.1f M warp-instr, %d samples" % (tot / 1e6, tots))
    ops = collections.Counter()
    for r in rs:
        ops[opname(r[1])] += int(r[iE])
    print("  mix:", ", ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in ops.most_common(18)))
    start = 0
    for i in range(1, len(rs) + 1):
        if i == len(rs) or abs(int(rs[i][iE]) - int(rs[start][iE])) > 0.15 * max(int(rs[start][iE]), 1):
            seg = rs[start:i]
            s = sum(int(r[iE]) for r in seg)
            if s > tot * minshare / 100.0:
                sm = sum(int(r[iS]) for r in seg)
                wf = sum(int(r[iW] or 0) for r in seg)
                wfi = sum(int(r[iWI] or 0) for r in seg)
                thr = sum(float(r[iT]) * int(r[iE]) for r in seg) / max(s, 1)
                mix = collections.Counter(opname(r[1]) for r in seg)
                print("  [%4d-%4d) n=%3d exec=%.2fM share=%4.1f%% samples=%4.1f%% smem_wf=%.1fM (ideal %.1fM) lanes=%.1f  %s" % (
                    start, i, i - start, s / (i - start) / 1e6, 100.0 * s / tot, 100.0 * sm / max(tots, 1), wf / 1e6, wfi / 1e6, thr,
                    dict(mix.most_common(7))))
            start = i
