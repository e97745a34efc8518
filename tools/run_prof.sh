#!/bin/bash
# Round-end measurement run on one GPU: full bench line, reference arm, ncu launch list of bench.py, ncu metrics of every kernel, C5 at spec.
M="gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-configs --frame-spp 0 --no-cpu-baseline > gpurun_out/r02_launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --frame-spp 0 --no-cpu-baseline > gpurun_out/r02_launch_ncu.log 2>&1; echo "launches rc=$?"
python tools/profile_all_kernels.py > gpurun_out/r02_all_plain.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_all_kernels.csv python tools/profile_all_kernels.py > gpurun_out/r02_all_ncu.log 2>&1; echo "all rc=$?"; tail -1 gpurun_out/r02_all_plain.log
for S in 1000000 10000000; do timeout 600 python bench.py --scene synthetic:$S --frame-spp 1024 --no-configs --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_${S}_1gpu.json 2> gpurun_out/r02_c5_${S}_1gpu.err; echo "c5 $S rc=$?"; done
python - <<PY
import json
for f in ["r02_bench_final", "r02_c5_1000000_1gpu", "r02_c5_10000000_1gpu"]:
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][0]); fr = d.get("frame") or {}
    print(f, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "frame", fr.get("spp"), "%.3f s" % fr.get("handle", {}).get("wall_s", float("nan")), "pairs/ray %.1f" % (d["roofline"]["per_ray"]["aabb_tests"] / 2), "bvh build ms", d.get("configs") and d["configs"].get("C5_1M", {}).get("bvh_build_ms"))
PY
