#!/usr/bin/env python3
"""ncu target for the kernels tools/profile_all_kernels.py reaches last (kept separate so that each ncu run stays short): the flat
extend kernel and the media-free fused kernel (Cornell), the media pass (Cornell-volume), the wide walk (sphere field), the
multi-GPU resolve kernel (one rank) and the two peak micro-benchmarks."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

kw = dict(num_samples=10000, max_depth=3, frames_per_batch=16, seed=1)
for name in ("cornell_original_test", "cornell_volume_10000_samples", "final_render_book_1"):
    tr = rt.RayTracer(rt.Scene.load(f"data/{name}.json"), **kw)
    tr.Update(16)
    tr.synchronize()
    if name == "cornell_original_test":
        tr.resolve_peers([tr.accum_ipc_handle()], 0, 16)
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
print(f"profile_rest_kernels done: fp32 peak {fp32:.1f} TFLOP/s, L2 read {v.value:.0f} GB/s")
