#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + SASS instruction counts between barriers."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "inst_executed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            print("%-70s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    for i, h in enumerate(hdr):
        if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
            try:
                if float(r[i]) >= 3.0:
                    print("%-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall "), r[i]))
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
except StopIteration:
    sys.exit(0)
hdr = rows[h]; data = []
for r in rows[h + 1:]:                      # a report with several kernels repeats the header: summarise the first kernel only
    if r == hdr:
        break
    if len(r) == len(hdr):
        data.append(r)
ia, ii, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ii]) for r in data); cur = 0; cs = 0; start = 0
print("total warp instructions", tot)
ops = {}
for k, r in enumerate(data):
    cur += int(r[ii]); cs += int(r[isamp])
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ii])
    if "BAR.SYNC" in r[ia] or k == len(data) - 1:
        print("SASS %4d-%4d inst %12d (%5.1f%%) samples %6d" % (start, k, cur, 100.0 * cur / max(tot, 1), cs)); cur = 0; cs = 0; start = k + 1
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
