#!/bin/bash
# The drop-in binary on all GPUs of the box: settings.json with num_samples 10000 (bounded: 40 s).
N=${1:-8}
R=$(mktemp -d); mkdir -p $R/local/data; ln -s $PWD/data $R/data
echo '{"num_samples": 10000, "render_once": true, "save_after_render_once": true, "max_depth": 50, "render_window": false}' > $R/local/data/settings.json
( cd $R && RAYTRACE2_ROOT=$R timeout 40 $OLDPWD/raytrace2_b200/bin/raytrace_2 data/book2_final_scene_10000_samples $OLDPWD/gpurun_out/r02_book2_10k_${N}gpu.png ) > gpurun_out/r02_raytrace2_${N}gpu.log 2>&1; echo "raytrace_2 rc=$?"; tail -2 gpurun_out/r02_raytrace2_${N}gpu.log
