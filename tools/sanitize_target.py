#!/usr/bin/env python3
"""Small target for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel family once, on tiny inputs.

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py

Covers: fixed-ray intersect (split, inline, flat, LBVH, wide), one render per pipeline variant (fused split / fused inline /
per-bin shade / flat extend / media pass / deferred noise shading / image textures off), the device LBVH build, the resolve
kernels (single rank, peer pointers through a multi-GPU handle when 2 GPUs are visible), texture hook, checkpoint write-back."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytrace2_b200 as rt  # noqa: E402
from _rays import fixed_rays  # noqa: E402


def scene(name):
    return rt.Scene.load(os.path.join(ROOT, "data", name + ".json"), perlin_seed=3)


CHECKS = {"enabled": None, "violations": {}, "nan_pixels": 0}


def collect(tr, img=None):
    """Fold the device-side self-check counters of `tr` (DEBUG_CHECKS build) and NaN pixels of `img` into CHECKS."""
    en, c = tr.debug_counters()
    CHECKS["enabled"] = en
    for k, v in c.items():
        CHECKS["violations"][k] = max(CHECKS["violations"].get(k, 0), v)  # the counters are per process, cumulative
    if img is not None:
        CHECKS["nan_pixels"] += int(np.isnan(img).sum())


def main():
    n_dev = rt.load_library().rt2_device_count()
    done = []
    book2 = scene("book2_final_scene_10000_samples")
    o, d, t = fixed_rays(book2, 3000, seed=1)
    for label, flags in [("unified", 0), ("split", rt.RT2_FLAG_INSTANCE_SPLIT), ("inline", rt.RT2_FLAG_INSTANCES_INLINE), ("lbvh", rt.RT2_FLAG_GPU_LBVH),
                         ("per-bin", rt.RT2_FLAG_NO_FUSED_SHADE), ("fast-math", rt.RT2_FLAG_FAST_MATH)]:
        tr = rt.RayTracer(book2, num_samples=4, max_depth=12, seed=2, flags=flags | rt.RT2_FLAG_MOMENTS, dims=(48, 48), frames_per_batch=2)
        tr.intersect(o, d, t)
        for _ in range(3):
            tr.Update(1)
        tr.NonConvertedPixels()
        tr.Pixels()
        s, ss = tr.read_accum(moments=True)
        tr.write_accum(s, ss, 3)
        tr.Update(1)
        tr.set_profiling(True)
        tr.Update(2)
        st = tr.stats()
        assert st["frames"] == 6 and st["stack_overflows"] == 0
        tr.texture_value(0, np.random.default_rng(0).uniform(-5, 5, (64, 3)).astype(np.float32))
        tr.OnResize((40, 24))
        tr.Update(1)
        collect(tr, tr.NonConvertedPixels())
        done.append(f"book2/{label}")
        del tr
    for name, flags in [("cornell_original_test", 0), ("cornell_original_test", rt.RT2_FLAG_NO_FLAT_EXTEND),
                        ("cornell_volume_10000_samples", 0), ("cornell_box4", rt.RT2_FLAG_NO_FLAT_EXTEND | rt.RT2_FLAG_GPU_LBVH),
                        ("final_render_book_1", 0), ("light_scene1", 0), ("checker_test", rt.RT2_FLAG_NO_FUSED_SHADE)]:
        sc = scene(name)
        tr = rt.RayTracer(sc, num_samples=4, max_depth=10, seed=5, flags=flags, dims=(40, 30), frames_per_batch=2)
        oo, dd, tt = fixed_rays(sc, 1500, seed=2)
        tr.intersect(oo, dd, tt)
        tr.Update(3)
        collect(tr, tr.NonConvertedPixels())
        handle = tr.accum_ipc_handle()
        tr.resolve_peers([handle], 0, 3)
        done.append(f"{name}/{flags}")
        del tr
    syn = rt.Scene.synthetic_spheres(6000, seed=9, width=48, height=27, host_bvh=False)
    for flags in (rt.RT2_FLAG_GPU_LBVH, rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_WIDE_BVH):
        tr = rt.RayTracer(syn, num_samples=2, max_depth=8, flags=flags, frames_per_batch=2)
        tr.Update(2)
        collect(tr, tr.NonConvertedPixels())
        done.append(f"synthetic/{flags}")
        del tr
    if n_dev >= 2:
        tr = rt.RayTracer(book2, num_samples=8, max_depth=10, seed=2, dims=(48, 48), frames_per_batch=2, n_gpus=2, flags=rt.RT2_FLAG_MOMENTS)
        tr.Update(5)
        collect(tr, tr.NonConvertedPixels())
        tr.Pixels()
        tr.read_accum(moments=True)
        done.append("multi-gpu handle")
    print("sanitize_target OK:", ", ".join(done))
    import json
    print("DEBUG_CHECKS " + json.dumps(CHECKS))


if __name__ == "__main__":
    main()
