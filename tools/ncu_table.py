#!/usr/bin/env python3
"""Markdown table from an `ncu --metrics ... --csv --log-file` capture: one row per kernel (launches merged: time summed,
ratios averaged weighted by time).   usage: tools/ncu_table.py all_kernels.csv > table.md"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(lambda: collections.defaultdict(list))  # kernel -> launch id -> {metric: value}
launch = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("rt2dev::", "").replace("rt2::", "")
    v = r[col["Metric Value"]].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        continue
    u = r[col["Metric Unit"]]
    if u in ("ns", "nsecond"):
        v *= 1e-3
    elif u in ("ms", "msecond"):
        v *= 1e3
    elif u == "Kbyte":
        v *= 1e3
    elif u == "Mbyte":
        v *= 1e6
    elif u == "Gbyte":
        v *= 1e9
    launch[(name, r[0])][r[col["Metric Name"]]] = v
agg = collections.defaultdict(list)
for (name, _), m in launch.items():
    agg[name].append(m)
S = {"t": "gpu__time_duration.sum", "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "l1hit": "l1tex__t_sector_hit_rate.pct",
     "l2hit": "lts__t_sector_hit_rate.pct", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tp": "l1tex__throughput.avg.pct_of_peak_sustained_active",
     "l2tp": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}
print("| kernel | launches | time µs (sum) | regs | lanes / instr | issue % | warps % | ALU / FMA / LSU pipe % | L1 throughput % | L1 hit % | L2 hit % | DRAM % of peak | DRAM GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, ms in sorted(agg.items(), key=lambda kv: -sum(m.get(S["t"], 0) for m in kv[1])):
    T = sum(m.get(S["t"], 0) for m in ms)
    if T <= 0:
        continue
    w = lambda k: sum(m.get(S[k], 0) * m.get(S["t"], 0) for m in ms) / T
    gbs = sum(m.get(S["rd"], 0) + m.get(S["wr"], 0) for m in ms) / (T * 1e-6) * 1e-9
    print(f"| `{name[:70]}` | {len(ms)} | {T:.0f} | {ms[0].get(S['regs'], 0):.0f} | {w('lanes'):.1f} | {w('issue'):.0f} | {w('warps'):.0f} | {w('alu'):.0f} / {w('fma'):.0f} / {w('lsu'):.0f} | "
          f"{w('l1tp'):.0f} | {w('l1hit'):.0f} | {w('l2hit'):.0f} | {w('dram'):.0f} | {gbs:.0f} |")
