#!/usr/bin/env python3
"""Small, deterministic target for ncu: one wavefront batch of the book-2 scene (or --scene) at its authored size."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="book2_final_scene_10000_samples")
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--fast-math", action="store_true")
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
scene = rt.Scene.load(f"data/{a.scene}.json")
tr = rt.RayTracer(scene, num_samples=10000, frames_per_batch=a.spp, seed=1, flags=(rt.RT2_FLAG_FAST_MATH if a.fast_math else 0) | a.flags)
tr.Update(a.spp)
st = tr.stats()
print(f"{a.scene}: {st['rays']} rays {st['gpu_ms_total']:.2f} ms {st['rays']/st['gpu_ms_total']*1e-3:.1f} Mrays/s launches {st['launches']}")
print("queue_sizes " + " ".join(str(int(x)) for x in tr.queue_sizes()))  # rays per bounce: launch b of the extend kernel traces [b]
