#!/bin/bash
# Short multi-GPU evidence run (N GPUs of one box, every command bounded): the raytrace_2 drop-in on all GPUs and on one, then
# bench.py under torchrun (weak-scaling value, e2e, both 10 000-spp frame legs; no configs table).
N=${1:-8}
nvidia-smi -L | head -$N
R=$(mktemp -d); mkdir -p $R/local/data; ln -s $PWD/data $R/data
echo '{"num_samples": 10000, "render_once": true, "save_after_render_once": true, "max_depth": 50, "render_window": false}' > $R/local/data/settings.json
( cd $R && RAYTRACE2_ROOT=$R timeout 40 $OLDPWD/raytrace2_b200/bin/raytrace_2 data/book2_final_scene_10000_samples $OLDPWD/gpurun_out/r02_book2_10k_${N}gpu.png ) > gpurun_out/r02_raytrace2_${N}gpu.log 2>&1; echo "raytrace_2 rc=$?"; tail -2 gpurun_out/r02_raytrace2_${N}gpu.log
( cd $R && RAYTRACE2_ROOT=$R timeout 40 $OLDPWD/raytrace2_b200/bin/raytrace_2 data/book2_final_scene_10000_samples /tmp/one.png --gpus 1 ) > gpurun_out/r02_raytrace2_1of${N}gpu.log 2>&1; tail -1 gpurun_out/r02_raytrace2_1of${N}gpu.log
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 --no-configs > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_${N}gpu.json") if l.startswith("{")][0]); fr = d.get("frame") or {}
    print("value %.0f e2e %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "| frame ranks %.3f s" % fr.get("ranks", {}).get("wall_s", float("nan")), "handle %.3f s" % fr.get("handle", {}).get("wall_s", float("nan")))
except Exception as e:
    print("FAILED", e)
PY
