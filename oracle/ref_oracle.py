"""ctypes binding of oracle/_ref/libref_oracle.so — the UNMODIFIED reference compiled by oracle/Makefile (`make ref`).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref_oracle.so")

_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{REF_LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_LIB_PATH)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_float)]
        L.ref_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float] + [C.c_void_p] * 6
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.ref_tracer_init.argtypes = [C.c_void_p, C.c_int]
        L.ref_tracer_update.argtypes = [C.c_void_p, C.c_int]
        L.ref_tracer_reset.argtypes = [C.c_void_p]
        L.ref_tracer_frame_idx.restype = C.c_uint64
        L.ref_tracer_frame_idx.argtypes = [C.c_void_p]
        L.ref_tracer_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_write_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_perlin_point_count.argtypes = [C.c_void_p, C.c_int]
        L.ref_perlin_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ref_perlin_set.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ref_texture_value.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_bvh_span1.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_top_level_is_medium.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """App.cpp:115-130 set-up of the reference on one scene file."""

    def __init__(self, path: str, num_samples: int = 1, dims=None):
        self.L = lib()
        w, h = dims if dims else (0, 0)
        self.h = self.L.ref_load(os.fsencode(path), num_samples, w, h)
        if not self.h:
            raise RuntimeError(f"reference loader failed on {path}")
        self.h = C.c_void_p(self.h)
        v = [C.c_int() for _ in range(5)]
        bg = (C.c_float * 3)()
        self.L.ref_info(self.h, *[C.byref(x) for x in v], bg)
        self.width, self.height, self.n_materials, self.n_textures, self.n_top = (x.value for x in v)
        self.background = np.array(list(bg), np.float32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_free(self.h)
            self.h = None

    def camera(self) -> np.ndarray:
        out = np.zeros(20, np.float32)
        self.L.ref_camera(self.h, _p(out))
        return out

    def intersect(self, origins, directions, times=None, tmin=0.001, tmax=3.402823466e+38):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros((n, 7), np.float32)
        rays[:, 0:3], rays[:, 3:6] = o, d
        if times is not None:
            rays[:, 6] = times
        out = {"hit": np.zeros(n, np.uint8), "t": np.zeros(n, np.float32), "point": np.zeros((n, 3), np.float32),
               "normal": np.zeros((n, 3), np.float32), "front_face": np.zeros(n, np.uint8), "material": np.zeros(n, np.int32)}
        self.L.ref_intersect(self.h, _p(rays), n, tmin, tmax, _p(out["hit"]), _p(out["t"]), _p(out["point"]), _p(out["normal"]),
                             _p(out["front_face"]), _p(out["material"]))
        return out

    def render(self, frame0: int, nframes: int, max_depth: int = 50, threads: int = 0, moments: bool = True):
        """Returns (sum, sumsq, n_rays, seconds); sum/sumsq are float64 [H, W, 3], row 0 = bottom."""
        if threads <= 0:
            threads = os.cpu_count() or 1
        s = np.zeros((self.height, self.width, 3), np.float64)
        ss = np.zeros_like(s) if moments else None
        nr, sec = C.c_uint64(), C.c_double()
        self.L.ref_render(self.h, frame0, nframes, max_depth, threads, _p(s), _p(ss) if moments else None, C.byref(nr), C.byref(sec))
        return s, ss, nr.value, sec.value

    # the real RayTracer object, serial
    def tracer_init(self, max_depth=50):
        self.L.ref_tracer_init(self.h, max_depth)

    def tracer_update(self, n=1):
        self.L.ref_tracer_update(self.h, n)

    def tracer_reset(self):
        self.L.ref_tracer_reset(self.h)

    def tracer_frame_idx(self) -> int:
        return self.L.ref_tracer_frame_idx(self.h)

    def tracer_read(self):
        mean = np.zeros((self.height, self.width, 3), np.float32)
        rgba = np.zeros((self.height, self.width, 4), np.uint8)
        self.L.ref_tracer_read(self.h, _p(mean), _p(rgba))
        return mean, rgba

    def perlin_get(self, tex_idx: int):
        n = self.L.ref_perlin_point_count(self.h, tex_idx)
        if n < 0:
            return None
        px, py, pz = (np.zeros(n, np.int32) for _ in range(3))
        vec = np.zeros((n, 3), np.float32)
        self.L.ref_perlin_get(self.h, tex_idx, _p(px), _p(py), _p(pz), _p(vec))
        return px, py, pz, vec

    def perlin_set(self, tex_idx: int, px, py, pz, vec):
        px, py, pz = (np.ascontiguousarray(a, np.int32) for a in (px, py, pz))
        vec = np.ascontiguousarray(vec, np.float32)
        return self.L.ref_perlin_set(self.h, tex_idx, _p(px), _p(py), _p(pz), _p(vec))

    def texture_value(self, tex_idx: int, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros_like(pts)
        self.L.ref_texture_value(self.h, tex_idx, _p(pts), pts.shape[0], _p(out))
        return out

    def span1_flags(self):
        f = np.zeros(max(self.n_top, 1), np.uint8)
        self.L.ref_bvh_span1(self.h, _p(f))
        return f[:self.n_top]


def write_image(rgb, path: str, png: bool = True):
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w, _ = rgb.shape
    lib().ref_write_image(_p(rgb), w, h, os.fsencode(path), int(png))
