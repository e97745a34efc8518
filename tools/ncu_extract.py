#!/usr/bin/env python3
"""DRAM traffic per ray of the extend kernel, from one `ncu --set full` capture of tools/profile_target.py:

    python tools/profile_target.py --spp 32 > plain.log && ncu --set full ... -k regex:k_traverse -s <first> -c <n> -o rep python tools/profile_target.py --spp 32
    python tools/ncu_extract.py rep.ncu-rep plain.log <first> profiles/r02_extend_traffic.json

The plain run prints the queue size of every bounce (rt2_read_queue_sizes); launch `first + i` of the extend kernel traces
bounce `first + i` (one extend launch per bounce in the unified / inline walks), so bytes per ray = (dram read + dram write) /
rays of that bounce.  bench.py multiplies the mean with its own rays per launch for `roofline.traffic`."""
import csv, io, json, subprocess, sys

rep, plain, first, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
sizes = [int(x) for x in [l for l in open(plain) if l.startswith("queue_sizes ")][-1].split()[1:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


launches = []
for k, r in enumerate([r for r in rows[2:] if len(r) >= len(hdr) and "k_traverse" in r[col["Kernel Name"]]]):
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    rays = sizes[first + k]
    launches.append({"bounce": first + k, "kernel": r[col["Kernel Name"]][:60], "rays": rays, "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "bytes_per_ray": (rd + wr) / rays, "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")),
                     "lanes_per_instruction": float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                     "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                     "l1_hit_pct": float(r[col["l1tex__t_sector_hit_rate.pct"]]), "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]])})
res = {"source": f"{rep} (ncu --set full of tools/profile_target.py, extend launches {first}..{first + len(launches) - 1})",
       "dram_bytes_per_ray": sum(l["bytes_per_ray"] * l["rays"] for l in launches) / sum(l["rays"] for l in launches),
       "algorithmic_bytes_per_ray": 48, "launches": launches}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "launches"}))
