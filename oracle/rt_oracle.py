"""oracle/rt_oracle.py — TEST INFRASTRUCTURE: Python front end of the CPU restatement (oracle/rt_oracle.cpp).

`PortScene(path)` restates `serialize::SceneLoader::LoadScene` (src/Serialize.cpp:199-360) + the set-up of `App::Run`
(src/App.cpp:115-130) in Python, building the object graph through the C calls of librt_oracle.so, whose classes mirror
the reference's Hittable tree one-to-one.  Legacy-format files (which HEAD's loader throws on) are adapted the same way the
product does (every primitive a top-level node) so configs like final_render_book_1 have an oracle too.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB_PATH = os.path.join(_HERE, "librt_oracle.so")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_LIB_PATH):
            build()
        L = C.CDLL(PORT_LIB_PATH)
        P, F3 = C.c_void_p, C.POINTER(C.c_float)
        L.orc_scene_new.restype = P
        L.orc_scene_free.argtypes = [P]
        L.orc_add_texture.argtypes = [P, C.c_int, F3, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_add_material.argtypes = [P, C.c_int, F3, C.c_float, C.c_float, C.c_int]
        L.orc_make_sphere.argtypes = [P, F3, F3, C.c_float, C.c_int]
        L.orc_make_quad.argtypes = [P, F3, F3, F3, C.c_int]
        L.orc_make_box.argtypes = [P, F3, F3, C.c_int]
        L.orc_make_medium.argtypes = [P, C.c_int, C.c_float, C.c_int]
        L.orc_make_list.argtypes = [P]
        L.orc_list_add.argtypes = [P, C.c_int, C.c_int]
        L.orc_make_transformed.argtypes = [P, C.c_int, F3, C.c_int, F3, F3]
        L.orc_add_top_level.argtypes = [P, C.c_int]
        L.orc_set_background.argtypes = [P, F3]
        L.orc_set_camera.argtypes = [P, F3, F3, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_finalize.argtypes = [P]
        L.orc_span1.argtypes = [P, P]
        L.orc_camera.argtypes = [P, P]
        L.orc_perlin_get.argtypes = [P, C.c_int, P, P, P, P]
        L.orc_perlin_set.argtypes = [P, C.c_int, P, P, P, P]
        L.orc_texture_value.argtypes = [P, C.c_int, P, C.c_size_t, P]
        L.orc_add_image_texture.argtypes = [P, C.c_int, C.c_int, P]
        L.orc_texture_value_uv.argtypes = [P, C.c_int, P, P, C.c_size_t, P]
        L.orc_intersect_uv.argtypes = [P, P, C.c_size_t, C.c_float, C.c_float, P, P]
        L.orc_intersect.argtypes = [P, P, C.c_size_t, C.c_float, C.c_float] + [P] * 7
        L.orc_render.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, P, P, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _f3(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(obj, key, default, n=3):
    v = obj.get(key, default) if isinstance(obj, dict) else default
    if not isinstance(v, list) or len(v) < n:
        v = default
    return [float(np.float32(x)) for x in v[:n]]


def _int_default(obj, key, default):
    """nlohmann value(key, <int default>): a JSON number is converted to int (truncation)."""
    v = obj.get(key, default)
    return int(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else default


MAT = {"lambertian": 0, "metal": 1, "dielectric": 2, "texture": 3, "diffuse_light": 4, "isotropic": 5}


def _load_image_linear(path):
    """[H, W, 3] float32, linear light, row 0 = top; a missing file is the book's 1x1 cyan debugging texture."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            a = np.asarray(im.convert("RGB"), np.uint8)
    except Exception:
        return np.array([[[0.0, 1.0, 1.0]]], np.float32)
    lut = (np.arange(256, dtype=np.float32) / np.float32(255.0)) ** np.float32(2.2)
    return np.ascontiguousarray(lut[a].astype(np.float32))


class PortScene:
    def __init__(self, path: str, num_samples: int = 1, dims=None, data_dir=None):
        self.L = lib()
        with open(path) as f:
            doc = json.load(f)
        self.h = C.c_void_p(self.L.orc_scene_new())
        L, h = self.L, self.h
        data_dir = data_dir or os.path.dirname(os.path.abspath(path))
        base = os.path.basename(path)

        # camera, Serialize.cpp:203-211, 32-40
        cam = doc.get("camera")
        if isinstance(cam, str):
            with open(os.path.join(data_dir, cam + ".json")) as f:
                camdoc = json.load(f)
        elif isinstance(cam, dict):
            camdoc = cam
        elif base.startswith("final_render"):
            with open(os.path.join(data_dir, "cam1.json")) as f:
                camdoc = json.load(f)
        else:
            camdoc = {}
        fov = float(_int_default(camdoc, "fov", 90))
        center = _arr(camdoc, "center", [0, 0, 1])
        look_at = _arr(camdoc, "look_at", [0, 0, 0])
        defocus = float(np.float32(camdoc.get("defocus_angle", 0.0)))
        focus = float(np.float32(camdoc.get("focus_distance", 1.0)))
        L.orc_set_background(h, _f3(_arr(doc, "background_color", [1, 1, 1])))

        self.n_textures = 0
        self.n_materials = 0
        self.noise_textures = []

        def add_tex(type_, albedo=(1, 1, 1), scale=1.0, even=0, odd=0, noise_type=1, point_count=256):
            idx = L.orc_add_texture(h, type_, _f3(albedo), scale, even, odd, noise_type, point_count)
            self.n_textures = idx + 1
            if type_ == 2:
                self.noise_textures.append(idx)
            return idx

        def add_mat(type_, albedo=(0, 0, 0), fuzz=0.0, ior=1.0, tex=0):
            idx = L.orc_add_material(h, type_, _f3(albedo), fuzz, ior, tex)
            self.n_materials = idx + 1
            return idx

        # textures, Serialize.cpp:216-242
        texs = doc.get("textures")
        if isinstance(texs, list):
            for t in texs:
                ty = t.get("type", "")
                if ty == "solid_color":
                    add_tex(0, _arr(t, "albedo", [1, 1, 1]))
                elif ty == "checker":
                    add_tex(1, scale=float(np.float32(t.get("scale", 1.0))), even=int(t.get("even_tex_idx", 0)), odd=int(t.get("odd_tex_idx", 0)))
                elif ty == "noise":
                    add_tex(2, _arr(t, "albedo", [1, 1, 1]), float(np.float32(t.get("scale", 1.0))),
                            noise_type=_int_default(t, "noise_type", 1), point_count=_int_default(t, "point_count", 256))
                elif ty == "image":
                    # schema extension: decoded with PIL (independent of the product's decoder), bytes -> linear by the 2.2 power law
                    rgb = _load_image_linear(os.path.join(data_dir, t.get("path", "")))
                    idx = L.orc_add_image_texture(h, rgb.shape[1], rgb.shape[0], _p(rgb))
                    self.n_textures = idx + 1
                else:
                    add_tex(0, [0, 0, 0])
        legacy = isinstance(doc.get("primitives"), dict)
        # materials, Serialize.cpp:244-285
        for m in doc["materials"]:
            ty = m.get("type", "")
            if not ty:
                if legacy and "tex_idx" in m:
                    ty = "texture"
                else:
                    raise ValueError("material type field empty")
            if ty == "lambertian":
                add_mat(0, _arr(m, "albedo", [1, 1, 1]))
            elif ty == "dielectric":
                add_mat(2, ior=float(np.float32(m.get("refraction_index", 1.0))))
            elif ty == "metal":
                add_mat(1, _arr(m, "albedo", [1, 1, 1]), fuzz=float(np.float32(m.get("fuzz", 0.0))))
            elif ty in ("texture", "diffuse_light"):
                code = 3 if ty == "texture" else 4
                if "tex_idx" in m:
                    add_mat(code, tex=int(m["tex_idx"]))
                elif "albedo" in m:
                    add_mat(code, tex=add_tex(0, _arr(m, "albedo", [1, 1, 1])))
                else:
                    add_mat(1)
            else:
                add_mat(1)

        def medium_wrap(p, handle):
            # Serialize.cpp:320-340
            cm = p.get("constant_medium")
            if cm is None:
                return handle
            if "albedo" in cm:
                mat = add_mat(5, tex=add_tex(0, _arr(cm, "albedo", [0, 0, 0])))
            elif "material" in cm:
                mat = int(cm["material"])
            else:
                return None
            return L.orc_make_medium(h, handle, float(np.float32(cm.get("density", 0.01))), mat)

        prims = []
        if not legacy:
            # Serialize.cpp:287-342
            for p in doc.get("primitives", []):
                ty = p.get("type", "")
                mat = _int_default(p, "material", 0)
                if ty == "quad":
                    hd = L.orc_make_quad(h, _f3(_arr(p, "q", [0, 0, 0])), _f3(_arr(p, "u", [1, 0, 0])), _f3(_arr(p, "v", [0, 0, 1])), mat)
                elif ty == "box":
                    hd = L.orc_make_box(h, _f3(_arr(p, "a", [0, 0, 0])), _f3(_arr(p, "b", [1, 1, 1])), mat)
                elif ty == "sphere":
                    hd = L.orc_make_sphere(h, _f3(_arr(p, "center", [0, 0, 0])), _f3(_arr(p, "displacement", [0, 0, 0])),
                                           float(np.float32(p.get("radius", 0.5))), mat)
                else:
                    continue
                hd = medium_wrap(p, hd)
                if hd is None:
                    continue
                prims.append(hd)
            for node in doc.get("scene", []):
                L.orc_add_top_level(h, self._node(node, prims))
        else:
            pr = doc["primitives"]
            ids = {m.get("id", i): i for i, m in enumerate(doc["materials"])}

            def mat_of(p):
                mid = _int_default(p, "material_id", _int_default(p, "material", 0))
                return ids.get(mid, mid)
            for p in pr.get("spheres", []):
                hd = medium_wrap(p, L.orc_make_sphere(h, _f3(_arr(p, "center", [0, 0, 0])), _f3(_arr(p, "displacement", [0, 0, 0])),
                                                      float(np.float32(p.get("radius", 0.5))), mat_of(p)))
                if hd is not None:
                    prims.append(hd)
            for p in pr.get("quads", []):
                hd = medium_wrap(p, L.orc_make_quad(h, _f3(_arr(p, "q", [0, 0, 0])), _f3(_arr(p, "u", [1, 0, 0])), _f3(_arr(p, "v", [0, 0, 1])), mat_of(p)))
                if hd is not None:
                    prims.append(hd)
            for p in pr.get("boxes", []):
                hd = medium_wrap(p, L.orc_make_box(h, _f3(_arr(p, "a", [0, 0, 0])), _f3(_arr(p, "b", [1, 1, 1])), mat_of(p)))
                if hd is not None:
                    prims.append(hd)
            for hd in prims:
                L.orc_add_top_level(h, hd)
        self.n_top = len(doc.get("scene", [])) if not legacy else len(prims)

        # dims, Serialize.cpp:349-357 + App.cpp:115,122-125
        w, hgt = 1600, 900
        if isinstance(cam, dict):
            width = _int_default(cam, "width", 0)
            aspect = float(np.float32(cam.get("aspect_ratio", 0.0)))
            if width != 0 and aspect != 0.0:
                hh = int(np.float32(width) / np.float32(aspect))
                if hh != 0:
                    w, hgt = width, hh
        if dims:
            w, hgt = dims
        self.width, self.height = w, hgt
        L.orc_set_camera(h, _f3(center), _f3(look_at), fov, defocus, focus, w, hgt, num_samples)
        L.orc_finalize(h)

    def _node(self, node, prims):
        # ParseNode, Serialize.cpp:161-197
        L, h = self.L, self.h
        obj = None
        if "primitive" in node:
            obj = prims[_int_default(node, "primitive", -1)]
        ch = node.get("children")
        if isinstance(ch, list):
            lst = L.orc_make_list(h)
            if obj is not None:
                L.orc_list_add(h, lst, obj)
            for c in ch:
                L.orc_list_add(h, lst, self._node(c, prims))
            obj = lst
        if obj is None:
            raise ValueError("error parsing node")
        tr = node.get("transform")
        if isinstance(tr, dict):
            has_rot = "rotation" in tr
            aa = _arr(tr, "rotation", [0, 0, 1, 0], 4) if has_rot else [0, 0, 1, 0]
            obj = L.orc_make_transformed(h, obj, _f3(_arr(tr, "translation", [0, 0, 0])), int(has_rot), _f3(aa), _f3(_arr(tr, "scale", [1, 1, 1])))
        return obj

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_scene_free(self.h)
            self.h = None

    def camera(self):
        out = np.zeros(18, np.float32)
        self.L.orc_camera(self.h, _p(out))
        return out

    def span1_flags(self):
        f = np.zeros(max(self.n_top, 1), np.uint8)
        self.L.orc_span1(self.h, _p(f))
        return f[:self.n_top]

    def intersect(self, origins, directions, times=None, tmin=0.001, tmax=3.402823466e+38):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros((n, 7), np.float32)
        rays[:, 0:3], rays[:, 3:6] = o, d
        if times is not None:
            rays[:, 6] = times
        out = {"hit": np.zeros(n, np.uint8), "t": np.zeros(n, np.float32), "point": np.zeros((n, 3), np.float32),
               "normal": np.zeros((n, 3), np.float32), "front_face": np.zeros(n, np.uint8), "material": np.zeros(n, np.int32),
               "leaf": np.zeros(n, np.int32)}
        self.L.orc_intersect(self.h, _p(rays), n, tmin, tmax, _p(out["hit"]), _p(out["t"]), _p(out["point"]), _p(out["normal"]),
                             _p(out["front_face"]), _p(out["material"]), _p(out["leaf"]))
        return out

    def render(self, frame0: int, nframes: int, max_depth: int = 50, threads: int = 0, moments: bool = True):
        if threads <= 0:
            threads = os.cpu_count() or 1
        s = np.zeros((self.height, self.width, 3), np.float64)
        ss = np.zeros_like(s) if moments else None
        nr, sec = C.c_uint64(), C.c_double()
        self.L.orc_render(self.h, frame0, nframes, max_depth, threads, _p(s), _p(ss) if moments else None, C.byref(nr), C.byref(sec))
        return s, ss, nr.value, sec.value

    def perlin_get(self, tex_idx: int):
        px, py, pz = (np.zeros(256, np.int32) for _ in range(3))
        vec = np.zeros((256, 3), np.float32)
        n = self.L.orc_perlin_get(self.h, tex_idx, _p(px), _p(py), _p(pz), _p(vec))
        return px[:n], py[:n], pz[:n], vec[:n]

    def perlin_set(self, tex_idx: int, px, py, pz, vec):
        px, py, pz = (np.ascontiguousarray(a, np.int32) for a in (px, py, pz))
        vec = np.ascontiguousarray(vec, np.float32)
        return self.L.orc_perlin_set(self.h, tex_idx, _p(px), _p(py), _p(pz), _p(vec))

    def texture_value_uv(self, tex_idx: int, pts, uv):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros_like(pts)
        self.L.orc_texture_value_uv(self.h, tex_idx, _p(pts), _p(uv), pts.shape[0], _p(out))
        return out

    def intersect_uv(self, origins, directions, times=None, tmin=0.001, tmax=3.402823466e+38):
        """(hit flags, HitRecord::uv) of the closest hit — Sphere.cpp:34,39-43 / Quad.cpp:15."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros((n, 7), np.float32)
        rays[:, 0:3], rays[:, 3:6] = o, d
        if times is not None:
            rays[:, 6] = times
        hit, uv = np.zeros(n, np.uint8), np.zeros((n, 2), np.float32)
        self.L.orc_intersect_uv(self.h, _p(rays), n, tmin, tmax, _p(hit), _p(uv))
        return hit, uv

    def texture_value(self, tex_idx: int, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.zeros_like(pts)
        self.L.orc_texture_value(self.h, tex_idx, _p(pts), pts.shape[0], _p(out))
        return out
