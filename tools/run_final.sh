#!/bin/bash
# Round-end verification + measurement on one GPU (every command bounded): GPU test suite, smoke, the full bench line, the ncu
# launch list of bench.py and one `ncu --set full` capture of the extend + fused kernels at bounces 2..3 of a 32-spp batch.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; tail -2 gpurun_out/r02_pytest_final.log
timeout 90 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 500 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
B="bench.py --steps 2 --warmup 3 --no-configs --frame-spp 0 --no-cpu-baseline"
timeout 100 python $B > gpurun_out/r02_launch_plain.log 2>&1 && timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/r02_launches.csv python $B > gpurun_out/r02_launch_ncu.log 2>&1; echo "launches rc=$?"
timeout 60 python tools/profile_target.py --spp 32 > gpurun_out/r02q_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_finish_shade" -s 4 -c 4 -f -o gpurun_out/r02q_b2b3 python tools/profile_target.py --spp 32 > gpurun_out/r02q_ncu.log 2>&1; echo "full rc=$?"; tail -2 gpurun_out/r02q_plain.log
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][0]); fr = d.get("frame") or {}
print("value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "frame %.3f s" % fr.get("handle", {}).get("wall_s", float("nan")), "roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "share_of_step")}, "cpu", d.get("cpu_baseline", {}).get("value"))
print({k: round(v.get("Mrays_per_s", 0)) for k, v in d["configs"].items()})
PY
