#!/bin/bash
# Re-tune of the while-while walk's round length (node steps per round) and refill threshold (busy lanes below which the warp
# fetches new rays) with the EXPERIMENTS build.  Usage: tools/run_tune_ab.sh "steps,fetch[,scene]" ...   (scene: book2 | book1 | c5 | vol)
# Every run is bounded (tools/ab_bench.sh: timeout 120 s) so that a bad setting cannot eat the GPU call.
mkdir -p gpurun_out
E=$PWD/raytrace2_b200/lib/libraytrace2_b200_exp.so
for sfx in "$@"; do
  IFS=, read s f scene <<< "$sfx"
  case "$scene" in
    book1) extra="--scene final_render_book_1";;
    c5) extra="--scene synthetic:1000000 --width 3840 --height 2160 --steps 3";;
    vol) extra="--scene cornell_volume_10000_samples";;
    *) extra=""; scene=book2;;
  esac
  tools/ab_bench.sh ${scene}_s${s}_f${f} RT2_LIB_PATH=$E RT2_TRAV_STEPS=$s RT2_TRAV_FETCH=$f -- $extra
done
