#!/usr/bin/env python
"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` (stdin or file)."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
stall = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio", h)]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for w in want:
        if w in d: print(f"{w:75s} {d[w]:>20s} {units[hdr.index(w)]}")
    st = sorted(((float(d[h]), h) for h in stall if d[h] not in ("", "n/a")), reverse=True)[:7]
    for v, h in st: print(f"   stall {h.split('stalled_')[1].split('_per_issue')[0]:28s} {v:6.2f}")
