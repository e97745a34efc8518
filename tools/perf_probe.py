#!/usr/bin/env python3
"""Quick perf probe: Mrays/s + per-kernel split for a scene at several settings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

def run(name, spp_total, spp, fpb, flags=0, dims=None, reps=3, label=""):
    lbvh = bool(flags & rt.RT2_FLAG_GPU_LBVH)
    scene = rt.Scene.load(f"data/{name}.json") if not name.startswith("synthetic:") else rt.Scene.synthetic_spheres(int(name.split(":")[1]), width=dims[0], height=dims[1], host_bvh=not lbvh)
    tr = rt.RayTracer(scene, num_samples=spp_total, frames_per_batch=fpb, flags=flags, seed=1, dims=dims)
    tr.Update(spp); tr.synchronize()
    tr.Reset()
    for _ in range(reps):
        tr.Update(spp)
    st = tr.stats()
    tr.Reset(); tr.set_profiling(True); tr.Update(spp); ps = tr.stats()
    print(f"[bvh build {ps['gpu_ms_bvh_build']:.2f} ms] " if ps['gpu_ms_bvh_build'] else "", end="")
    print(f"{label or name}: spp_total={spp_total} spp={spp} fpb={fpb} flags={flags} dims={tr.Dims()} -> {st['rays']/st['gpu_ms_total']*1e-3:.1f} Mrays/s "
          f"({st['gpu_ms_total']/reps:.2f} ms per {spp} spp, rays/path {st['rays']/st['paths']:.3f}) | profiled split ms: traverse {ps['gpu_ms_extend']:.2f} finish {ps['gpu_ms_finish']:.2f} shade {ps['gpu_ms_shade']:.2f} other {ps['gpu_ms_other']:.2f} | per ray: box-pairs {ps['box_pair_tests']/ps['rays']:.1f} spheres {ps['sphere_tests']/ps['rays']:.1f} quads {ps['quad_tests']/ps['rays']:.1f} inst {ps['instance_visits']/ps['rays']:.2f}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "book2"
    if which == "book2":
        n = "book2_final_scene_10000_samples"
        run(n, 64, 64, 0)
        run(n, 10000, 64, 0)
        run(n, 10000, 16, 16)
        run(n, 64, 16, 16)
        run(n, 64, 64, 0, flags=rt.RT2_FLAG_FAST_MATH)
    elif which == "synthetic":
        for n in (1000000,):
            run(f"synthetic:{n}", 1024, 4, 0, dims=(1920, 1080), label=f"synthetic {n} SAH")
            run(f"synthetic:{n}", 1024, 4, 0, dims=(1920, 1080), flags=rt.RT2_FLAG_GPU_LBVH, label=f"synthetic {n} LBVH")
        run("synthetic:10000000", 1024, 2, 0, dims=(1920, 1080), flags=rt.RT2_FLAG_GPU_LBVH, label="synthetic 10M LBVH")
    elif which == "lbvh":
        for n in ["cornell_original_test", "book2_final_scene_10000_samples", "final_render_book_1"]:
            run(n, 10000, 32, 0, flags=rt.RT2_FLAG_GPU_LBVH, label=n + " [LBVH]")
    elif which == "all":
        for n in ["cornell_original_test", "cornell_volume_10000_samples", "book2_final_scene_10000_samples", "final_render_book_1"]:
            run(n, 10000, 32, 0)
            run(n, 10000, 32, 0, flags=rt.RT2_FLAG_FAST_MATH, label=n + " [fast]")
