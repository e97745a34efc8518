"""Multi-GPU: one process per GPU, samples (frames) partitioned across ranks, accumulators summed on rank 0.

The path shards naturally over (pixel, frame) because the RNG is counter-based: rank g of G traces the
global frames f ≡ g (mod G), which keeps every rank's strata spread over the sqrt(spp) x sqrt(spp) grid
(s_i = f % sqrt, s_j = f / sqrt % sqrt, RayTracer.cpp:59-60).  The only exchange step is one float32 sum-reduce of the
W*H accumulators at read-out time: `torch.distributed.reduce` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def frame_partition(total_frames: int, rank: int, world_size: int) -> Tuple[int, int, int]:
    """(offset, stride, local_count): rank traces global frames offset + k*stride, k < local_count."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    local = (total_frames - rank + world_size - 1) // world_size if total_frames > rank else 0
    return rank, world_size, local


def frames_of_rank(total_frames: int, rank: int, world_size: int) -> List[int]:
    off, stride, n = frame_partition(total_frames, rank, world_size)
    return [off + k * stride for k in range(n)]


class _CudaBuffer:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__`` v2)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class DistributedRayTracer:
    """Wraps one local tracer per rank.  `tracer` needs Update / FrameIdx / Dims / read_accum (and accum_device_ptr for
    the NCCL path); `raytrace2_b200.RayTracer` provides them, the CPU tests pass a stub."""

    def __init__(self, tracer, total_frames: int, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.tracer = tracer
        self.total_frames = total_frames
        self.offset, self.stride, self.local_frames = frame_partition(total_frames, self.rank, self.world_size)

    def render(self) -> None:
        """Trace this rank's share of the frames (no communication)."""
        done = self.tracer.FrameIdx()
        if self.local_frames > done:
            self.tracer.Update(self.local_frames - done)

    def reduce_accum(self) -> Optional[np.ndarray]:
        """Sum of all ranks' accumulators on rank 0 (float32 [H, W, 3]); None elsewhere."""
        import torch
        dist = self.dist
        backend = dist.get_backend(self.group)
        w, h = self.tracer.Dims()
        if backend == "nccl":
            ptr, n = self.tracer.accum_device_ptr()
            self.tracer.synchronize()
            acc = torch.as_tensor(_CudaBuffer(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
            # Reduce a COPY: the renderer's accumulator must keep holding this rank's frames only, or a second read-out, a later
            # Update + read-out, resolve_p2p() or a checkpoint on rank 0 would count the other ranks' frames twice.
            t = acc.clone()
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM, group=self.group)
            torch.cuda.synchronize()
            if self.rank != 0:
                return None
            return t.view(h, w, 4)[..., :3].cpu().numpy()
        acc = torch.from_numpy(np.ascontiguousarray(self.tracer.read_accum()))
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        return acc.numpy() if self.rank == 0 else None

    def resolve_p2p(self, rgba8: bool = False) -> Optional[np.ndarray]:
        """Read-out without a collective: rank 0 sums the peers' accumulators straight out of their HBM (CUDA IPC + NVLink
        P2P loads) inside the resolve kernel (`rt2_resolve_peers`), in rank order — reduce + mean (+ RGBA8) in one pass.
        One node, NCCL process group (the handles travel through `all_gather_object`)."""
        dist = self.dist
        handles = [None] * self.world_size
        dist.all_gather_object(handles, self.tracer.accum_ipc_handle(), group=self.group)
        self.tracer.synchronize()
        dist.barrier(group=self.group)       # every rank has finished rendering
        out = None
        if self.rank == 0:
            out = self.tracer.resolve_peers(handles, 0, self.total_frames, rgba8=rgba8)
        dist.barrier(group=self.group)       # peers may touch their accumulators again
        return out

    def NonConvertedPixels(self) -> Optional[np.ndarray]:
        """Mean over ALL ranks' frames on rank 0 (≡ RayTracer::NonConvertedPixels after `total_frames` Updates)."""
        total = self.reduce_accum()
        if total is None:
            return None
        return total / np.float32(self.total_frames)
