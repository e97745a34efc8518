#!/usr/bin/env python3
"""ncu target that launches EVERY kernel of the shipped library at a realistic size: the unified / split / inline instance walks,
the fused finish + shade kernels (simple-media, media-free and general variants), deferred noise shading, the media pass, the flat
and wide extend kernels, the per-bin pipeline, the device LBVH build (Karras and PLOC), generate / accumulate / resolve, the
texture hook and the peak micro-benchmarks.  Book 2 at 600 x 600, 16 spp, 4 bounces unless noted.

    python tools/profile_all_kernels.py > plain.log && ncu --metrics <list> --csv --log-file all.csv python tools/profile_all_kernels.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

kw = dict(num_samples=10000, max_depth=4, frames_per_batch=16, seed=1)
book2 = rt.Scene.load("data/book2_final_scene_10000_samples.json")
for flags in (0, rt.RT2_FLAG_INSTANCE_SPLIT, rt.RT2_FLAG_INSTANCES_INLINE, rt.RT2_FLAG_NO_FUSED_SHADE, rt.RT2_FLAG_GPU_LBVH,
              rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_LBVH_PLOC, rt.RT2_FLAG_FAST_MATH):
    tr = rt.RayTracer(book2, flags=flags, **kw)
    tr.Update(16)
    tr.NonConvertedPixels()
    tr.Pixels()
    if flags == 0:
        tr.texture_value(0, np.zeros((4096, 3), np.float32))
    del tr
for name in ("cornell_original_test", "cornell_volume_10000_samples", "final_render_book_1", "light_scene1"):
    tr = rt.RayTracer(rt.Scene.load(f"data/{name}.json"), **kw)
    tr.Update(16)
    tr.synchronize()
    del tr
field = rt.Scene.synthetic_spheres(200000, width=1280, height=720, host_bvh=False)
for flags in (rt.RT2_FLAG_GPU_LBVH, rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_WIDE_BVH):
    tr = rt.RayTracer(field, num_samples=1024, max_depth=3, frames_per_batch=2, seed=1, flags=flags)
    tr.Update(2)
    tr.synchronize()
    del tr
lib = rt.load_library()
v = C.c_double()
lib.rt2_measure_fp32_peak(0, C.byref(v))
fp32 = v.value
lib.rt2_measure_l2_bandwidth(0, C.byref(v))
print(f"profile_all_kernels done: fp32 peak {fp32:.1f} TFLOP/s, L2 read {v.value:.0f} GB/s")
