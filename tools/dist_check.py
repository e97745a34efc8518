#!/usr/bin/env python3
"""Multi-GPU correctness on real GPUs (run under torchrun, NCCL): the image of N ranks — read out (a) by one kernel on rank 0
over peer memory (rt2_resolve_peers: CUDA IPC + NVLink P2P loads) and (b) by an NCCL sum-reduce of the accumulators — equals the
single-GPU render of the same global frames up to fp32 summation order."""
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
# read-out 1: one kernel on rank 0 that reads the peers' accumulators over NVLink (no collective); does not modify them
import time
drt.resolve_p2p()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
p2p = drt.resolve_p2p()
t_p2p = time.perf_counter() - t0
p2p8 = drt.resolve_p2p(rgba8=True)
# read-out 2: NCCL sum-reduce into rank 0's accumulator, then the mean
t0 = time.perf_counter()
img = drt.NonConvertedPixels()
t_nccl = time.perf_counter() - t0
# the NCCL read-out reduces a copy: a second read-out and a peer-memory read-out afterwards must see the same accumulators
img_again = drt.NonConvertedPixels()
p2p_again = drt.resolve_p2p()
if rank == 0:
    assert np.array_equal(img, img_again), "a second NCCL read-out must not double-count rank 0's frames"
    assert np.array_equal(p2p, p2p_again), "the NCCL read-out must leave the renderer's accumulator untouched"
if rank == 0:
    d = np.abs(p2p - img).max() / max(img.max(), 1e-9)
    print(f"dist_check: P2P resolve vs NCCL reduce: max rel diff {d:.3e}; wall {t_p2p * 1e3:.2f} ms vs {t_nccl * 1e3:.2f} ms (incl. barriers / host copy)")
    assert d < 1e-5, "the peer-memory read-out must equal the NCCL read-out up to fp32 reassociation"
    want8 = np.floor(np.clip(p2p.astype(np.float64), 0, 1) * 255.999).astype(np.uint8)
    assert np.array_equal(p2p8[..., :3], want8) and np.all(p2p8[..., 3] == 255)
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
