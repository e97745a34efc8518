#!/usr/bin/env python3
"""ncu target that launches EVERY kernel of the library once or a few times at a realistic size: device LBVH build, generate,
traverse, ray sort, fused finish+shade, deferred shade kernels, accumulate, resolve (book 2, 600x600, 16 spp, 3 bounces),
the per-bin pipeline (k_finish_hit + k_shade_scatter<*>), the flat extend kernel (Cornell) and the wide walk (sphere field)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

os.environ["RT2_SORT_MIN"] = "1000"
book2 = rt.Scene.load("data/book2_final_scene_10000_samples.json")
tr = rt.RayTracer(book2, num_samples=10000, max_depth=3, frames_per_batch=16, seed=1, flags=rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_SORT_RAYS)
tr.Update(16)
tr.NonConvertedPixels()
tr.Pixels()
del tr
tr = rt.RayTracer(book2, num_samples=10000, max_depth=3, frames_per_batch=16, seed=1, flags=rt.RT2_FLAG_NO_FUSED_SHADE)
tr.Update(16)
tr.synchronize()
del tr
cornell = rt.Scene.load("data/cornell_original_test.json")
tr = rt.RayTracer(cornell, num_samples=10000, max_depth=3, frames_per_batch=16, seed=1)
tr.Update(16)
tr.synchronize()
del tr
field = rt.Scene.synthetic_spheres(200000, width=1280, height=720, host_bvh=False)
tr = rt.RayTracer(field, num_samples=1024, max_depth=3, frames_per_batch=2, seed=1, flags=rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_WIDE_BVH)
tr.Update(2)
tr.synchronize()
print("profile_all_kernels done")
