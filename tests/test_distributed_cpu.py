"""Multi-GPU host logic on CPU: world_size-2 gloo processes, frame partition + accumulator reduce."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT
from raytrace2_b200.distributed import frame_partition, frames_of_rank


def test_frame_partition_covers_every_frame_once():
    for total in (0, 1, 7, 16, 100, 10000):
        for world in (1, 2, 3, 4, 8):
            allf = sorted(f for r in range(world) for f in frames_of_rank(total, r, world))
            assert allf == list(range(total))
            counts = [frame_partition(total, r, world)[2] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        frame_partition(10, 2, 2)


def test_partition_keeps_strata_spread():
    """f = rank (mod G) visits every stratum column s_i = f % sqrt(spp) on every rank when gcd(G, sqrt) == 1, and G/gcd of
    them otherwise — never a single column (RayTracer.cpp:59-60)."""
    spp, sq = 10000, 100
    for world in (2, 4, 8):
        for r in range(world):
            cols = {f % sq for f in frames_of_rank(spp, r, world)}
            rows = {(f // sq) % sq for f in frames_of_rank(spp, r, world)}
            assert len(cols) == sq // np.gcd(world, sq) and len(rows) == sq


class StubTracer:
    """Stands in for raytrace2_b200.RayTracer: frame f contributes the image (f+1) * pattern."""

    def __init__(self, offset, stride, w=8, h=4):
        self.offset, self.stride, self.w, self.h = offset, stride, w, h
        self.frames = 0
        self.pattern = np.arange(w * h * 3, dtype=np.float32).reshape(h, w, 3) / 7.0
        self.acc = np.zeros_like(self.pattern)

    def Update(self, n):
        for _ in range(n):
            f = self.offset + self.frames * self.stride
            self.acc += np.float32(f + 1) * self.pattern
            self.frames += 1

    def FrameIdx(self):
        return self.frames

    def Dims(self):
        return self.w, self.h

    def read_accum(self):
        return self.acc.copy()


def _worker(rank, world, total, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from raytrace2_b200.distributed import DistributedRayTracer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, stride, _ = frame_partition(total, rank, world)
    drt = DistributedRayTracer(StubTracer(off, stride), total)
    drt.render()
    drt.render()  # idempotent: no extra frames
    img = drt.NonConvertedPixels()
    if rank == 0:
        np.save(os.path.join(out_dir, "mean.npy"), img)
    else:
        assert img is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 16])
def test_two_rank_gloo_reduce_equals_single_process(tmp_path, total):
    world = 2
    port = 29500 + (os.getpid() % 2000) + total
    mp.spawn(_worker, args=(world, total, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "mean.npy")
    single = StubTracer(0, 1)
    single.Update(total)
    want = single.acc / np.float32(total)
    assert np.allclose(got, want, rtol=1e-6)
