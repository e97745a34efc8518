#!/bin/bash
# usage: tools/exp_all.sh "<env A>" "<env B>" ...   (runs the 3 repo scenes + synthetic 1M under each setting)
B=book2_final_scene_10000_samples
python tools/exp_probe.py $B --spp 64 -- "$@"
python tools/exp_probe.py cornell_original_test --spp 64 -- "$@"
python tools/exp_probe.py final_render_book_1 --spp 16 -- "$@"
python tools/exp_probe.py synthetic:1000000 --dims 1920x1080 --spp 8 -- "$@"
