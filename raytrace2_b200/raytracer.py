"""Host-side mirror of the reference's renderer interface over the C ABI (include/rt2.h).

Reference interfaces mirrored (paths relative to the reference root):
  * ``serialize::SceneLoader::LoadScene``  src/Serialize.hpp:21-22   -> :class:`SceneLoader`, :class:`Scene`
  * ``raytrace2::cpu::RayTracer``           src/cpu_raytrace/RayTracer.hpp:15-42 -> :class:`RayTracer`
  * ``util::WriteImage``                    src/Util.hpp:11-12        -> :func:`WriteImage`
  * ``App::Run`` (headless branch)          src/App.cpp:81-249        -> :func:`run_app`
Method names keep the reference's spelling so the parity tests read like the reference's call sites (src/App.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import check, load_library

HIT_DTYPE = np.dtype([("point", np.float32, 3), ("t", np.float32), ("normal", np.float32, 3), ("material", np.int32),
                      ("prim", np.uint32), ("instance", np.int32), ("front_face", np.uint32), ("uv16", np.uint32)])
assert HIT_DTYPE.itemsize == C.sizeof(_capi.Hit)


def _as_np(ptr, count, struct):
    """Copy `count` structs at `ptr` into a numpy structured array."""
    if count == 0:
        return np.zeros(0, dtype=np.dtype(struct))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(struct * count)).contents).copy()


class Scene:
    """A compiled scene: the flattened SoA buffers + BVH the device consumes (≡ cpu::Scene after App.cpp:122-126)."""

    def __init__(self, handle: int):
        self._lib = load_library()
        self._h = C.c_void_p(handle)

    # ---- construction ----
    @classmethod
    def load(cls, path: str, data_dir: Optional[str] = None, perlin_seed: int = 0) -> "Scene":
        lib = load_library()
        h = C.c_void_p()
        check(lib.rt2_scene_load(os.fsencode(path), os.fsencode(data_dir) if data_dir else None, perlin_seed, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_builder(cls, builder, data_dir: Optional[str] = None, perlin_seed: int = 0) -> "Scene":
        """Compiles a :class:`raytrace2_b200.scene_builder.SceneBuilder` document."""
        return cls.from_string(builder.to_json(), data_dir, perlin_seed)

    @classmethod
    def from_string(cls, text: str, data_dir: Optional[str] = None, perlin_seed: int = 0) -> "Scene":
        lib = load_library()
        h = C.c_void_p()
        check(lib.rt2_scene_load_string(text.encode(), os.fsencode(data_dir) if data_dir else None, perlin_seed, C.byref(h)))
        return cls(h.value)

    @classmethod
    def synthetic_spheres(cls, n: int, seed: int = 20261018, width: int = 3840, height: int = 2160, host_bvh: bool = True) -> "Scene":
        """BASELINE config 5.  host_bvh=False skips the host SAH build; render such a scene with RT2_FLAG_GPU_LBVH."""
        lib = load_library()
        h = C.c_void_p()
        check(lib.rt2_scene_synthetic_spheres(n, seed, width, height, int(host_bvh), C.byref(h)))
        return cls(h.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rt2_scene_destroy(h)

    # ---- inspection ----
    @property
    def desc(self) -> _capi.SceneDesc:
        d = _capi.SceneDesc()
        check(self._lib.rt2_scene_get_desc(self._h, C.byref(d)))
        return d

    @property
    def dims(self) -> Tuple[int, int]:
        d = self.desc
        return d.width, d.height

    def set_dims(self, width: int, height: int) -> None:
        check(self._lib.rt2_scene_set_dims(self._h, width, height))

    def spheres(self):
        d = self.desc
        return _as_np(d.spheres, d.n_spheres, _capi.Sphere)

    def quads(self):
        d = self.desc
        return _as_np(d.quads, d.n_quads, _capi.Quad)

    def xforms(self):
        d = self.desc
        return _as_np(d.xforms, d.n_xforms, _capi.Xform)

    def instances(self):
        d = self.desc
        return _as_np(d.instances, d.n_instances, _capi.Instance)

    def media(self):
        d = self.desc
        return _as_np(d.media, d.n_media, _capi.Medium)

    def materials(self):
        d = self.desc
        return _as_np(d.materials, d.n_materials, _capi.Material)

    def textures(self):
        d = self.desc
        return _as_np(d.textures, d.n_textures, _capi.Texture)

    def images(self):
        """List of [H, W, 4] float32 arrays (linear RGBA, row 0 = top), one per image texture."""
        d = self.desc
        out = []
        if d.n_images:
            tab = _as_np(d.images, d.n_images, _capi.Image)
            tex = np.ctypeslib.as_array(d.image_texels, shape=(d.n_image_texels, 4))
            for im in tab:
                o, w, h = int(im["texel_offset"]), int(im["width"]), int(im["height"])
                out.append(tex[o:o + w * h].reshape(h, w, 4).copy())
        return out

    def nodes(self):
        d = self.desc
        return _as_np(d.nodes, 2 * d.n_node_pairs, _capi.BvhNode)

    def prim_refs(self):
        d = self.desc
        if d.n_prim_refs == 0:
            return np.zeros(0, np.uint32)
        return np.ctypeslib.as_array(d.prim_refs, shape=(d.n_prim_refs,)).copy()

    def span1_flags(self) -> np.ndarray:
        n = self.desc.n_top_level
        out = np.zeros(max(n, 1), np.uint8)
        check(self._lib.rt2_scene_span1_flags(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[:n]

    def get_perlin(self, perlin_idx: int = 0):
        px, py, pz = (np.zeros(256, np.int32) for _ in range(3))
        vec = np.zeros((256, 3), np.float32)
        check(self._lib.rt2_scene_get_perlin(self._h, perlin_idx, px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                             pz.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p)))
        return px, py, pz, vec

    def set_perlin(self, perlin_idx: int, px, py, pz, vec) -> None:
        px, py, pz = (np.ascontiguousarray(a, np.int32) for a in (px, py, pz))
        vec = np.ascontiguousarray(vec, np.float32).reshape(256, 3)
        check(self._lib.rt2_scene_set_perlin(self._h, perlin_idx, px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p),
                                             pz.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p)))


class SceneLoader:
    """``serialize::SceneLoader`` (src/Serialize.hpp:21-31): ``LoadScene`` returns None where the reference returns nullopt."""

    def __init__(self, data_dir: Optional[str] = None, perlin_seed: int = 0):
        self.data_dir = data_dir
        self.perlin_seed = perlin_seed

    def LoadScene(self, filepath: str) -> Optional[Scene]:
        try:
            return Scene.load(filepath, self.data_dir, self.perlin_seed)
        except _capi.Rt2Error as e:
            # Serialize.cpp:102-104
            print(f"Failed to parse Scene: {e.message}. {filepath}", file=sys.stderr)
            return None


class RayTracer:
    """``raytrace2::cpu::RayTracer`` (src/cpu_raytrace/RayTracer.hpp:15-42) on the B200s of one box.

    ``num_samples`` is ``AppSettings::num_samples`` as App.cpp:129 hands it to ``Camera::SetSamplesPerPixel``: it fixes
    the stratification grid (sqrt(num_samples) cells per axis, RayTracer.cpp:57-60).  ``n_gpus`` > 1 (or -1 = all visible)
    puts one replica per GPU behind this object: the frames of every ``Update`` are dealt round-robin and every read-out
    sums the replicas over peer memory (rt2_config.n_gpus).  ``frame_offset`` / ``frame_stride`` select which global frames
    this instance traces when the partition is done OUTSIDE (one process per GPU, ``raytrace2_b200.distributed``).
    """

    def __init__(self, scene: Scene, *, num_samples: int = 1, max_depth: int = 50, device: int = 0, seed: int = 0x5EED,
                 flags: int = 0, frames_per_batch: int = 0, frame_offset: int = 0, frame_stride: int = 1,
                 dims: Optional[Tuple[int, int]] = None, n_gpus: int = 1):
        self._lib = load_library()
        self.scene = scene
        self.max_depth = max_depth
        cfg = _capi.Config()
        cfg.device = device
        cfg.width, cfg.height = dims if dims else (0, 0)
        cfg.samples_per_pixel = num_samples
        cfg.max_depth = max_depth
        cfg.frames_per_batch = frames_per_batch
        cfg.frame_offset = frame_offset
        cfg.frame_stride = frame_stride
        cfg.flags = flags
        cfg.seed = seed
        cfg.n_gpus = n_gpus
        self._cfg = cfg
        h = C.c_void_p()
        check(self._lib.rt2_create(scene._h, C.byref(cfg), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rt2_destroy(h)

    close = __del__

    # ---- the reference's public surface ----
    def Update(self, n_frames: int = 1) -> None:
        """RayTracer::Update (RayTracer.cpp:55-70), `n_frames` times: +1 sample per pixel each."""
        check(self._lib.rt2_update(self._h, n_frames))

    def OnResize(self, dims: Sequence[int]) -> None:
        check(self._lib.rt2_resize(self._h, int(dims[0]), int(dims[1])))

    def Reset(self) -> None:
        check(self._lib.rt2_reset(self._h))

    def FrameIdx(self) -> int:
        v = C.c_uint64()
        check(self._lib.rt2_frame_idx(self._h, C.byref(v)))
        return v.value

    def Dims(self) -> Tuple[int, int]:
        w, h = C.c_int32(), C.c_int32()
        check(self._lib.rt2_dims(self._h, C.byref(w), C.byref(h)))
        return w.value, h.value

    def NonConvertedPixels(self) -> np.ndarray:
        """Mean radiance, float32 [H, W, 3], row 0 = bottom of the image (RayTracer.cpp:105-112)."""
        w, h = self.Dims()
        out = np.empty((h, w, 3), np.float32)
        check(self._lib.rt2_read_mean_rgb32f(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def Pixels(self) -> np.ndarray:
        """RGBA8 preview [H, W, 4], linear, no gamma (RayTracer.cpp:16-18,65-66)."""
        w, h = self.Dims()
        out = np.empty((h, w, 4), np.uint8)
        check(self._lib.rt2_read_rgba8(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- extras (test / measurement hooks) ----
    def synchronize(self) -> None:
        check(self._lib.rt2_synchronize(self._h))

    def flush(self) -> None:
        """Trace every frame requested by ``Update`` so far (asynchronous; ``Update`` collects frames into wavefront batches)."""
        check(self._lib.rt2_flush(self._h))

    def texture_value(self, tex_idx: int, points, uv=None) -> np.ndarray:
        """Device ``textures[tex_idx]->Value(u, v, p)`` (Texture.cpp:7-22) for fixed points: [n, 3] float32."""
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        n = p.shape[0]
        out = np.zeros((n, 3), np.float32)
        uvp = None
        if uv is not None:
            uva = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
            assert uva.shape[0] == n
            uvp = uva.ctypes.data_as(C.c_void_p)
        check(self._lib.rt2_texture_value(self._h, tex_idx, p.ctypes.data_as(C.c_void_p), uvp, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def upload_scene(self, scene: Optional[Scene] = None) -> None:
        check(self._lib.rt2_upload_scene(self._h, (scene or self.scene)._h))

    def read_accum(self, moments: bool = False):
        w, h = self.Dims()
        s = np.empty((h, w, 3), np.float32)
        ss = np.empty((h, w, 3), np.float32) if moments else None
        check(self._lib.rt2_read_accum(self._h, s.ctypes.data_as(C.c_void_p), ss.ctypes.data_as(C.c_void_p) if moments else None))
        return (s, ss) if moments else s

    def write_accum(self, s: np.ndarray, ss: Optional[np.ndarray], frames: int) -> None:
        s = np.ascontiguousarray(s, np.float32)
        ss = None if ss is None else np.ascontiguousarray(ss, np.float32)
        check(self._lib.rt2_write_accum(self._h, s.ctypes.data_as(C.c_void_p), ss.ctypes.data_as(C.c_void_p) if ss is not None else None, frames))

    def save_checkpoint(self, path: str) -> None:
        """Accumulators + frame index + the settings a resume must match, as one .npz (the reference has no checkpointing:
        SURVEY §5).  Resume with :meth:`load_checkpoint` on a renderer created with the same scene / seed / num_samples."""
        moments = bool(self._cfg.flags & _capi.RT2_FLAG_MOMENTS)
        acc = self.read_accum(moments=moments)
        s, ss = acc if moments else (acc, None)
        w, h = self.Dims()
        np.savez(path, sum=s, sumsq=ss if ss is not None else np.zeros(0, np.float32), frames=np.uint64(self.FrameIdx()),
                 dims=np.array([w, h], np.int32), seed=np.uint64(self._cfg.seed), num_samples=np.int32(self._cfg.samples_per_pixel),
                 max_depth=np.int32(self._cfg.max_depth), frame_offset=np.int32(self._cfg.frame_offset), frame_stride=np.int32(self._cfg.frame_stride))

    def load_checkpoint(self, path: str) -> int:
        """Restores a checkpoint written by :meth:`save_checkpoint`; returns the number of frames it holds."""
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        w, h = self.Dims()
        if tuple(int(x) for x in z["dims"]) != (w, h):
            raise ValueError(f"checkpoint is {tuple(z['dims'])}, renderer is {(w, h)}")
        for key, mine in (("seed", self._cfg.seed), ("num_samples", self._cfg.samples_per_pixel), ("max_depth", self._cfg.max_depth),
                          ("frame_offset", self._cfg.frame_offset), ("frame_stride", max(1, self._cfg.frame_stride))):
            theirs = int(z[key]) if key != "frame_stride" else max(1, int(z[key]))
            if theirs != int(mine):
                raise ValueError(f"checkpoint {key}={theirs} does not match the renderer's {int(mine)}")
        ss = z["sumsq"] if z["sumsq"].size else None
        if ss is not None and not (self._cfg.flags & _capi.RT2_FLAG_MOMENTS):
            ss = None
        frames = int(z["frames"])
        self.write_accum(z["sum"], ss, frames)
        return frames

    def accum_device_ptr(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        check(self._lib.rt2_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_ipc_handle(self) -> bytes:
        """RT2_IPC_HANDLE_BYTES (80) bytes naming the accumulator for peers (multi-GPU read-out over peer memory, rt2_resolve_peers)."""
        buf = (C.c_uint8 * 80)()
        check(self._lib.rt2_accum_ipc_handle(self._h, buf))
        return bytes(buf)

    def resolve_peers(self, handles: Sequence[bytes], self_rank: int, total_frames: int, rgba8: bool = False) -> np.ndarray:
        """Mean (or RGBA8 preview) over the accumulators of all ranks, summed in rank order by ONE kernel that reads the
        peers' HBM over NVLink.  `handles[r]` = rank r's :meth:`accum_ipc_handle`.  The caller has synchronised all ranks."""
        w, h = self.Dims()
        blob = b"".join(bytes(x) for x in handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        if rgba8:
            out = np.empty((h, w, 4), np.uint8)
            check(self._lib.rt2_resolve_peers(self._h, arr, len(handles), self_rank, total_frames, None, out.ctypes.data_as(C.c_void_p)))
        else:
            out = np.empty((h, w, 3), np.float32)
            check(self._lib.rt2_resolve_peers(self._h, arr, len(handles), self_rank, total_frames, out.ctypes.data_as(C.c_void_p), None))
        return out

    def set_frame_idx(self, frames: int) -> None:
        check(self._lib.rt2_set_frame_idx(self._h, frames))

    def stream(self) -> int:
        s = C.c_void_p()
        check(self._lib.rt2_stream(self._h, C.byref(s)))
        return s.value or 0

    def set_profiling(self, enabled: bool) -> None:
        check(self._lib.rt2_set_profiling(self._h, int(enabled)))

    def queue_sizes(self) -> np.ndarray:
        """Rays traced at every bounce of the most recent wavefront batch (first GPU)."""
        out = np.zeros(max(self.max_depth, 1), np.uint32)
        n = C.c_uint32(0)
        check(self._lib.rt2_read_queue_sizes(self._h, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out[:n.value]

    def debug_counters(self):
        """(enabled, {check: violations}) of the device-side self checks (a `make DEBUG_CHECKS=1` build; else enabled is False)."""
        names = ["node", "sphere", "quad", "instance", "inst_leaf", "stack", "queue", "entry", "material", "texture", "slot", "bin",
                 "medium", "prim_ref", "nan", "reserved"]
        buf = (C.c_uint64 * 16)()
        en = C.c_int(0)
        check(self._lib.rt2_debug_counters(self._h, buf, C.byref(en)))
        return bool(en.value), {n: int(v) for n, v in zip(names, buf)}

    def stats(self) -> dict:
        st = _capi.Stats()
        check(self._lib.rt2_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _capi.Stats._fields_}

    def read_bvh(self):
        """(nodes, prim_refs, tlas_root) of the BVH the device traverses (host SAH upload or device LBVH build)."""
        np_, nr, root = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(self._lib.rt2_read_bvh(self._h, None, 0, None, 0, C.byref(np_), C.byref(nr), C.byref(root)))
        nodes = np.zeros(2 * np_.value, np.dtype(_capi.BvhNode))
        refs = np.zeros(max(nr.value, 1), np.uint32)
        check(self._lib.rt2_read_bvh(self._h, nodes.ctypes.data_as(C.c_void_p), nodes.size, refs.ctypes.data_as(C.c_void_p), refs.size,
                                     C.byref(np_), C.byref(nr), C.byref(root)))
        return nodes, refs[:nr.value], root.value

    def intersect(self, origins, directions, times=None, tmin: float = 0.001, tmax: float = 3.402823466e+38,
                  skip_media: bool = False) -> np.ndarray:
        """Fixed-ray closest hit ≡ scene.hittable_list.Hit(r, Interval{tmin, tmax}) (RayTracer.cpp:25)."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros((n, 8), np.float32)
        rays[:, 0:3] = o
        rays[:, 3] = 0.0 if times is None else np.asarray(times, np.float32)
        rays[:, 4:7] = d
        out = np.zeros(n, HIT_DTYPE)
        check(self._lib.rt2_intersect(self._h, rays.ctypes.data_as(C.c_void_p), n, tmin, tmax, int(skip_media),
                                      out.ctypes.data_as(C.c_void_p)))
        return out


def WriteImage(pixels: np.ndarray, width: int, height: int, out_path: str, png: bool = True) -> None:
    """``util::WriteImage(vector<vec3>, w, h, path, png)`` (src/Util.cpp:39-79): sqrt gamma, vertical flip, 8-bit RGB."""
    px = np.ascontiguousarray(pixels, np.float32).reshape(height, width, 3)
    check(load_library().rt2_write_image(px.ctypes.data_as(C.c_void_p), width, height, os.fsencode(out_path), int(png)))


def tonemap_rgb8(pixels: np.ndarray) -> np.ndarray:
    px = np.ascontiguousarray(pixels, np.float32)
    h, w, _ = px.shape
    out = np.empty((h, w, 3), np.uint8)
    check(load_library().rt2_tonemap_rgb8(px.ctypes.data_as(C.c_void_p), w, h, out.ctypes.data_as(C.c_void_p)))
    return out


def run_app(argv: Sequence[str], settings_path: str, data_dir: Optional[str] = None) -> int:
    """Headless ``App::Run`` (src/App.cpp:81-249): ``argv`` as for ``raytrace_2 [scene[.json]] [out.png]``."""
    lib = load_library()
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    return lib.rt2_app_run(len(argv), arr, os.fsencode(settings_path), os.fsencode(data_dir) if data_dir else None)
