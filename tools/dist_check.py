#!/usr/bin/env python3
"""Multi-GPU correctness on real GPUs (run under torchrun, NCCL): the reduced image of N ranks equals the single-GPU render of
the same global frames up to fp32 summation order."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytrace2_b200 as rt

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = 64
scene = rt.Scene.load(os.path.join(ROOT, "data", "cornell_volume_10000_samples.json"))
off, stride, n_local = rt.frame_partition(total, rank, world)
tr = rt.RayTracer(scene, num_samples=total, device=local, seed=5, frame_offset=off, frame_stride=stride, dims=(256, 256))
drt = rt.DistributedRayTracer(tr, total)
drt.render()
img = drt.NonConvertedPixels()
if rank == 0:
    single = rt.RayTracer(scene, num_samples=total, device=local, seed=5, dims=(256, 256))
    single.Update(total)
    want = single.NonConvertedPixels()
    err = np.abs(img - want).max()
    rel = err / max(want.max(), 1e-9)
    print(f"dist_check: world={world} max abs diff {err:.3e} (rel {rel:.3e}) mean {want.mean():.5f}")
    assert rel < 1e-5, "N-GPU result must equal the 1-GPU result up to fp32 reassociation"
    print("dist_check OK")
dist.barrier()
dist.destroy_process_group()
