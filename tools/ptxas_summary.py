#!/usr/bin/env python3
"""Summarise a `-Xptxas -v` log: registers, spills, stack, shared memory per kernel (demangled)."""
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "raytrace2_b200/build/ptxas_rt_kernels.log"
text = open(path).read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*?Used (\d+) registers(.*)", text):
    name, stack, st, ld, regs, rest = m.groups()
    smem = re.search(r"(\d+) bytes smem", rest)
    rows.append((name, int(regs), int(stack), int(st), int(ld), int(smem.group(1)) if smem else 0))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"{'regs':>4} {'stack':>5} {'sp_st':>5} {'sp_ld':>5} {'smem':>5}  kernel")
for (n, regs, stack, st, ld, smem), dn in zip(rows, names):
    dn = re.sub(r"\(.*", "", dn).replace("rt2dev::", "").replace("void ", "")
    print(f"{regs:4d} {stack:5d} {st:5d} {ld:5d} {smem:5d}  {dn}")
