"""raytrace2_b200 — B200-native (sm_100a) path-tracing backend, drop-in for tonadr1022/Raytrace2's ``src/cpu_raytrace``.

Python is only the host-side mirror of the reference's renderer interface (``RayTracer``: Update / OnResize / Reset /
Pixels / NonConvertedPixels / FrameIdx / Dims; ``SceneLoader.LoadScene``; ``WriteImage``) over the C ABI in
``include/rt2.h``.  All rendering happens in hand-written CUDA kernels inside ``lib/libraytrace2_b200.so``.
"""
from ._capi import (LIB_PATH, RT2_FLAG_FAST_MATH, RT2_FLAG_GPU_LBVH, RT2_FLAG_MOMENTS, RT2_FLAG_NO_FUSED_SHADE, RT2_FLAG_NO_FLAT_EXTEND, RT2_FLAG_INSTANCES_INLINE, RT2_FLAG_LBVH_PLOC, RT2_FLAG_FLOAT_NODES, RT2_FLAG_INSTANCE_SPLIT, RT2_FLAG_NO_INSTANCE_SPLIT, RT2_FLAG_SORT_RAYS, RT2_FLAG_WIDE_BVH, Rt2Error, load_library)
from .raytracer import RayTracer, Scene, SceneLoader, WriteImage, run_app
from . import scene_builder
from .distributed import DistributedRayTracer, frame_partition

__all__ = ["LIB_PATH", "RT2_FLAG_FAST_MATH", "RT2_FLAG_GPU_LBVH", "RT2_FLAG_MOMENTS", "RT2_FLAG_NO_FUSED_SHADE", "RT2_FLAG_NO_FLAT_EXTEND", "RT2_FLAG_INSTANCES_INLINE", "RT2_FLAG_LBVH_PLOC", "RT2_FLAG_FLOAT_NODES", "RT2_FLAG_INSTANCE_SPLIT", "RT2_FLAG_NO_INSTANCE_SPLIT", "RT2_FLAG_SORT_RAYS", "RT2_FLAG_WIDE_BVH", "Rt2Error", "load_library",
           "RayTracer", "Scene", "SceneLoader", "WriteImage", "run_app", "DistributedRayTracer", "frame_partition"]
