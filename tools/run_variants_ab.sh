#!/bin/bash
# A/B of variant builds of the library (make EXPERIMENTS=1 EXTRA_DEFS=... OUTLIB=..._<tag>.so) against the shipped one, book 2,
# plus the hit-parity tests under each variant.  Usage: tools/run_variants_ab.sh tag ...   (every run bounded)
mkdir -p gpurun_out
tools/ab_bench.sh ship_book2
for v in "$@"; do
  L=$PWD/raytrace2_b200/lib/libraytrace2_b200_$v.so
  tools/ab_bench.sh ${v}_book2 RT2_LIB_PATH=$L
  RT2_LIB_PATH=$L timeout 150 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_round2.py -m gpu -x -q -k "not binary and not debug" > gpurun_out/variant_pytest_$v.log 2>&1; tail -1 gpurun_out/variant_pytest_$v.log
done
