#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on): stall samples, instruction share and
average active lanes per line, for kernels whose name contains the given substring.
usage: tools/ncu_lines.py report.ncu-rep kernel_substring [min_pct]"""
import csv, subprocess, sys, io
rep, sub = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
recs, cur_file, hdr, fn = [], None, None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and sub in fn:
        d = dict(zip(hdr, r))
        try:
            recs.append((cur_file, int(r[0]), r[1].strip()[:95], int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
        except (KeyError, ValueError):
            pass
# merge duplicates (several launches)
agg = {}
for f, ln, src, s, i, t in recs:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0, 0])
    a[1] += s; a[2] += i; a[3] += t
ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values()); tt = sum(a[3] for a in agg.values())
print(f"kernel~{sub}: samples {ts} inst {ti} avg lanes {tt / max(ti, 1):.2f}")
for (f, ln), a in sorted(agg.items()):
    if a[1] * 100 >= min_pct * ts or a[2] * 100 >= min_pct * ti:
        print(f"{f:14s}:{ln:4d} samp {100 * a[1] / ts:5.1f}% inst {100 * a[2] / ti:5.1f}% lanes {a[3] / max(a[2], 1):5.1f} | {a[0]}")
