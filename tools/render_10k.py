#!/usr/bin/env python3
"""The north-star run: data/book2_final_scene_10000_samples.json at its authored 600x600, 10 000 spp, max_depth 50, samples
partitioned over all ranks (torchrun, one process per GPU, NCCL reduce), PNG written by rank 0.  Prints one JSON line:
wall seconds of render + reduce (and of everything including scene load and the PNG), rays, Mrays/s, paths/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/render_10k.py [--spp 10000] [--scene NAME]
"""
import argparse, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytrace2_b200 as rt

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="book2_final_scene_10000_samples")
ap.add_argument("--spp", type=int, default=10000)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "book2_10k.png"))
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t_start = time.perf_counter()
scene = rt.Scene.load(os.path.join(ROOT, "data", a.scene + ".json"))
off, stride, n_local = rt.frame_partition(a.spp, rank, world)
tr = rt.RayTracer(scene, num_samples=a.spp, max_depth=50, device=local, seed=20261018, frame_offset=off, frame_stride=stride)
tr.Update(8); tr.synchronize(); tr.Reset()          # warm-up (context, allocator), not counted
if world > 1:
    warm = torch.zeros(1, device="cuda"); dist.all_reduce(warm); torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
if world > 1:
    drt = rt.DistributedRayTracer(tr, a.spp)
    drt.render()
    img = drt.NonConvertedPixels()                    # NCCL sum-reduce to rank 0 + mean
else:
    tr.Update(a.spp)
    img = tr.NonConvertedPixels()
torch.cuda.synchronize()
t1 = time.perf_counter()
st = tr.stats()
tot = torch.tensor([float(st["rays"]), float(st["paths"]), st["gpu_ms_total"]], dtype=torch.float64, device="cuda")
mx = tot.clone()
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    w, h = tr.Dims()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    rt.WriteImage(img, w, h, a.out, True)
    t2 = time.perf_counter()
    rays, paths = float(tot[0]), float(tot[1])
    print(json.dumps({"scene": a.scene, "dims": [w, h], "spp": a.spp, "n_gpus": world, "render_plus_reduce_s": t1 - t0,
                      "total_s_incl_load_warmup_png": t2 - t_start, "device_ms_max_over_ranks": float(mx[2]),
                      "rays": rays, "paths": paths, "Mrays_per_s": rays / (t1 - t0) * 1e-6, "paths_per_s": paths / (t1 - t0),
                      "mean_radiance": float(np.asarray(img).mean()), "png": os.path.relpath(a.out, ROOT)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
