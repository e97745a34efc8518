#!/usr/bin/env python3
"""Experiment probe: Mrays/s + per-kernel split of a scene under several tuning settings (env RT2_* read at rt2_create).

    python tools/exp_probe.py <scene|synthetic:N> [--dims WxH] [--spp S] [--flags F] -- "RT2_SORT=1" "RT2_SORT=1 RT2_SORT_MIN=65536" ...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt


def run(name, spp, flags, dims, env, reps=3):
    for k in [k for k in os.environ if k.startswith("RT2_")]:
        del os.environ[k]
    for kv in env.split():
        k, v = kv.split("=")
        os.environ[k] = v
    lbvh = bool(flags & rt.RT2_FLAG_GPU_LBVH)
    if name.startswith("synthetic:"):
        scene = rt.Scene.synthetic_spheres(int(name.split(":")[1]), width=dims[0], height=dims[1], host_bvh=not lbvh)
    else:
        scene = rt.Scene.load(f"data/{name}.json")
    tr = rt.RayTracer(scene, num_samples=10000, frames_per_batch=0, flags=flags, seed=1, dims=dims)
    tr.Update(spp); tr.synchronize(); tr.Reset()
    for _ in range(reps):
        tr.Update(spp)
    st = tr.stats()
    tr.Reset(); tr.set_profiling(True); tr.Update(spp); ps = tr.stats()
    r = ps["rays"]
    print(f"{name} [{env or 'default'}] flags={flags} dims={tr.Dims()} spp={spp}: {st['rays']/st['gpu_ms_total']*1e-3:8.1f} Mrays/s "
          f"({st['gpu_ms_total']/reps:.2f} ms/step) | split ms: trav {ps['gpu_ms_extend']:.2f} sort {ps['gpu_ms_sort']:.2f} finish {ps['gpu_ms_finish']:.2f} "
          f"shade {ps['gpu_ms_shade']:.2f} other {ps['gpu_ms_other']:.2f} | /ray: box {ps['box_pair_tests']/r:.1f} sph {ps['sphere_tests']/r:.2f} "
          f"quad {ps['quad_tests']/r:.2f} inst {ps['instance_visits']/r:.2f}", flush=True)
    del tr, scene


if __name__ == "__main__":
    args = sys.argv[1:]
    envs = [""]
    if "--" in args:
        i = args.index("--")
        envs = args[i + 1:]
        args = args[:i]
    name = args[0]
    dims, spp, flags = None, 32, 0
    for j, a in enumerate(args):
        if a == "--dims":
            dims = tuple(int(x) for x in args[j + 1].split("x"))
        if a == "--spp":
            spp = int(args[j + 1])
        if a == "--flags":
            flags = int(args[j + 1])
    for e in envs:
        run(name, spp, flags, dims, e)
