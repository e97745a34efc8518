#!/bin/bash
# BVH builder sweep (RT2_BVH_MAX_LEAF / RT2_BVH_TRAV_COST are read once per process): one python process per setting
for L in 1 4; do
  export RT2_BVH_MAX_LEAF=$L
  echo "== MAX_LEAF=$L"
  python tools/exp_probe.py final_render_book_1 --spp 16 -- "RT2_BVH_MAX_LEAF=$L" 2>&1 | cut -c1-230
  python tools/exp_probe.py synthetic:1000000 --dims 1920x1080 --spp 8 -- "RT2_BVH_MAX_LEAF=$L" 2>&1 | cut -c1-230
  python tools/exp_probe.py cornell_box_scene_graph --spp 64 -- "RT2_BVH_MAX_LEAF=$L RT2_FLAT=0" 2>&1 | cut -c1-230
  python tools/exp_probe.py final_render_scene_blur --spp 16 -- "RT2_BVH_MAX_LEAF=$L" 2>&1 | cut -c1-230
done
