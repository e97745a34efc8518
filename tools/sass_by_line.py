#!/usr/bin/env python3
"""Static code size of one kernel by source line: SASS instructions (x16 bytes) attributed through nvdisasm -g line info.
usage: tools/sass_by_line.py <file.cubin|.o> <kernel substring> [top]   (a .o is unpacked with cuobjdump -xelf)"""
import collections, os, re, subprocess, sys, tempfile
src, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if src.endswith(".o"):
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(src)], cwd=td, capture_output=True)
    src = os.path.join(td, [f for f in os.listdir(td) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", src], capture_output=True, text=True).stdout.splitlines()
cur_fn, cur_line, counts, total = None, None, collections.Counter(), 0
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur_fn and sub in cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        counts[cur_line] += 1
        total += 1
print(f"{sub}: {total} SASS instructions = {total * 16} bytes")
for (f, l), c in counts.most_common(top):
    print(f"  {c * 16:6d} B  {f}:{l}")
