#!/bin/bash
# A/B of traversal variants on one box: shipped library against variant builds (EXTRA_DEFS), book 2 unless noted.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_round2.py tests/test_gpu_lbvh.py tests/test_gpu_pipeline.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/quant_pytest.log 2>&1; tail -3 gpurun_out/quant_pytest.log
for v in "$@"; do
  RT2_LIB_PATH=$PWD/raytrace2_b200/lib/libraytrace2_b200_$v.so timeout 300 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_round2.py -m gpu -x -q -k "not binary and not debug" > gpurun_out/quant_pytest_$v.log 2>&1; tail -1 gpurun_out/quant_pytest_$v.log
done
tools/ab_bench.sh q_book2
for v in "$@"; do tools/ab_bench.sh ${v}_book2 RT2_LIB_PATH=$PWD/raytrace2_b200/lib/libraytrace2_b200_$v.so; done
tools/ab_bench.sh q_book1 -- --scene final_render_book_1
tools/ab_bench.sh q_c5_1m -- --scene synthetic:1000000 --width 3840 --height 2160 --steps 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*_book*.json")+glob.glob("gpurun_out/ab_*c5*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), d["config"]["compact_nodes"], d["config"]["node_inflation"])
    except Exception as e: print(f, e)
PY
