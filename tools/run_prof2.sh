#!/bin/bash
# Second measurement run on one GPU: C5 at spec (3840 x 2160, 1024 spp) on 1 GPU, and the ncu metrics of the remaining kernels.
M="gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed"
for S in 1000000 10000000; do timeout 200 python bench.py --scene synthetic:$S --frame-spp 1024 --no-configs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_${S}_1gpu.json 2> gpurun_out/r02_c5_${S}_1gpu.err; echo "c5 $S rc=$?"; done
timeout 100 python tools/profile_rest_kernels.py > gpurun_out/r02_rest_plain.log 2>&1 && timeout 240 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_rest_kernels.csv python tools/profile_rest_kernels.py > gpurun_out/r02_rest_ncu.log 2>&1; echo "rest rc=$?"; tail -1 gpurun_out/r02_rest_plain.log
python - <<PY
import json
for f in ["r02_c5_1000000_1gpu", "r02_c5_10000000_1gpu"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][0]); fr = d.get("frame") or {}
        print(f, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "frame", fr.get("spp"), "%.3f s" % fr.get("handle", {}).get("wall_s", float("nan")), "pairs/ray %.1f" % (d["roofline"]["per_ray"]["aabb_tests"] / 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
