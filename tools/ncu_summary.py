#!/usr/bin/env python3
"""One line per kernel launch of an .ncu-rep (--set full): duration, issue utilisation, lanes per instruction, pipes, cache
hit rates, DRAM bytes, top stalls.   usage: tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
W = ["gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
     "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")] or \
         [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]][:70]
    print(f"== {name}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for w in W:
        if w in col:
            print(f"   {w:75s} {r[col[w]]}")
    st = sorted(((float(r[col[h]].replace(',', '')) if r[col[h]] not in ('', 'n/a') else 0.0, h) for h in stalls), reverse=True)[:6]
    for v, h in st:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per')[0]:30s} {v:.2f}")
