#!/bin/bash
# The multi-GPU evidence run (N GPUs of one box): GPU tests, bench.py under torchrun, the raytrace_2 drop-in on all GPUs, C5 at spec.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -$N
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/r02_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_${N}gpu.log
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
# the drop-in binary: settings.json with num_samples 10000, all GPUs of the box
R=$(mktemp -d); mkdir -p $R/local/data; ln -s $PWD/data $R/data
echo '{"num_samples": 10000, "render_once": true, "save_after_render_once": true, "max_depth": 50, "render_window": false}' > $R/local/data/settings.json
( cd $R && RAYTRACE2_ROOT=$R $OLDPWD/raytrace2_b200/bin/raytrace_2 data/book2_final_scene_10000_samples $OLDPWD/gpurun_out/r02_book2_10k_${N}gpu.png ) > gpurun_out/r02_raytrace2_${N}gpu.log 2>&1; echo "raytrace_2 rc=$?"; cat gpurun_out/r02_raytrace2_${N}gpu.log
( cd $R && RAYTRACE2_ROOT=$R $OLDPWD/raytrace2_b200/bin/raytrace_2 data/book2_final_scene_10000_samples /tmp/one.png --gpus 1 ) > gpurun_out/r02_raytrace2_1of${N}gpu.log 2>&1; tail -2 gpurun_out/r02_raytrace2_1of${N}gpu.log
# C5 at spec: 3840 x 2160, 1024 spp
for S in 1000000 10000000; do
  timeout 900 $TR --master-port 29522 bench.py --gpus $N --scene synthetic:$S --frame-spp 1024 --no-configs --steps 3 --warmup 3 > gpurun_out/r02_c5_${S}_${N}gpu.json 2> gpurun_out/r02_c5_${S}_${N}gpu.err; echo "c5 $S rc=$?"
done
python - <<PY
import json
for f in ["r02_bench_${N}gpu", "r02_c5_1000000_${N}gpu", "r02_c5_10000000_${N}gpu"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][0])
        fr = d.get("frame") or {}
        print(f, "value %.0f e2e %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
              "| frame spp", fr.get("spp"), "ranks %.3f s" % fr.get("ranks", {}).get("wall_s", float("nan")), "handle %.3f s" % fr.get("handle", {}).get("wall_s", float("nan")),
              "| pairs/ray %.1f" % (d["roofline"]["per_ray"]["aabb_tests"] / 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
