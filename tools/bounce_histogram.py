#!/usr/bin/env python3
"""Queue size per bounce: rays(max_depth = d) - rays(max_depth = d - 1) for d = 1..50 (same seed, so the paths are identical)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt
name = sys.argv[1] if len(sys.argv) > 1 else "book2_final_scene_10000_samples"
spp = 4
scene = rt.Scene.load(f"data/{name}.json")
prev, out, work, pw = 0, [], [], (0, 0, 0)
for d in range(1, 51):
    tr = rt.RayTracer(scene, num_samples=spp, max_depth=d, seed=1)
    tr.set_profiling(True)  # the counting build of k_traverse
    tr.Update(spp)
    st = tr.stats()
    r = st["rays"]
    w = (st["box_pair_tests"], st["sphere_tests"] + st["quad_tests"], st["instance_visits"])
    out.append(r - prev)
    work.append(tuple(a - b for a, b in zip(w, pw)))
    prev, pw = r, w
    del tr
n0 = out[0]
print(name, "paths", n0, "rays/path", prev / n0)
print("fraction of paths alive at bounce b:", " ".join(f"{b}:{out[b] / n0:.4f}" for b in (0, 1, 2, 3, 5, 8, 10, 12, 15, 20, 25, 30, 40, 49)))
print("share of all rays in bounces >= 12:", sum(out[12:]) / prev)
print("node pairs / primitive tests / instance visits per ray at bounce b:",
      " ".join(f"{b}:{work[b][0] / out[b]:.1f}/{work[b][1] / out[b]:.2f}/{work[b][2] / out[b]:.2f}" for b in (0, 1, 2, 3, 5, 8, 12, 20, 30, 49)))
