"""ctypes binding of the C ABI declared in include/rt2.h (libraytrace2_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C raytrace2_b200/csrc``.  There is no Python or
CPU fallback: if the shared object is missing, importing a symbol from here raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT2_LIB_PATH") or os.path.join(_HERE, "lib", "libraytrace2_b200.so")

RT2_OK = 0
RT2_ERR_INVALID_ARG = -1
RT2_ERR_IO = -2
RT2_ERR_PARSE = -3
RT2_ERR_CUDA = -4
RT2_ERR_UNSUPPORTED = -5
RT2_ERR_STATE = -6

RT2_FLAG_MOMENTS = 1
RT2_FLAG_FAST_MATH = 2
RT2_FLAG_NO_FUSED_SHADE = 4
RT2_FLAG_GPU_LBVH = 8
RT2_FLAG_SORT_RAYS = 16
RT2_FLAG_WIDE_BVH = 32
RT2_FLAG_INSTANCES_INLINE = 64
RT2_FLAG_NO_INSTANCE_SPLIT = 64
RT2_FLAG_INSTANCE_SPLIT = 256
RT2_FLAG_LBVH_PLOC = 512
RT2_FLAG_FLOAT_NODES = 1024
RT2_FLAG_NO_FLAT_EXTEND = 128
RT2_MAX_HOISTED_INSTANCES = 4
RT2_ABI_VERSION = 3

RT2_PRIM_SPHERE, RT2_PRIM_QUAD, RT2_PRIM_INSTANCE, RT2_PRIM_MEDIUM = 0, 1, 2, 3
RT2_PRIM_NONE = 0xFFFFFFFF

MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_TEXTURE, MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC = range(6)
TEX_SOLID, TEX_CHECKER, TEX_NOISE = range(3)


class Sphere(C.Structure):
    _fields_ = [("center0", C.c_float * 3), ("radius", C.c_float), ("displacement", C.c_float * 3), ("material", C.c_uint32)]


class Quad(C.Structure):
    _fields_ = [("normal", C.c_float * 3), ("d", C.c_float), ("q", C.c_float * 3), ("material", C.c_uint32),
                ("u", C.c_float * 3), ("pad0", C.c_float), ("v", C.c_float * 3), ("pad1", C.c_float),
                ("w", C.c_float * 3), ("pad2", C.c_float)]


class Xform(C.Structure):
    _fields_ = [("inv", (C.c_float * 4) * 3), ("model", (C.c_float * 4) * 3)]


class Instance(C.Structure):
    _fields_ = [("chain_first", C.c_uint32), ("chain_len", C.c_uint32), ("blas_root", C.c_uint32), ("top_level_node", C.c_uint32)]


class Medium(C.Structure):
    _fields_ = [("neg_inv_density", C.c_float), ("material", C.c_uint32), ("boundary_first", C.c_uint32),
                ("boundary_count", C.c_uint32), ("chain_first", C.c_uint32), ("chain_len", C.c_uint32),
                ("sample_twice", C.c_uint32), ("top_level_node", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("tex_idx", C.c_uint32), ("fuzz", C.c_float), ("refraction_index", C.c_float),
                ("albedo", C.c_float * 3), ("pad", C.c_float)]


class Texture(C.Structure):
    _fields_ = [("type", C.c_uint32), ("even_tex_idx", C.c_uint32), ("odd_tex_idx", C.c_uint32), ("perlin_idx", C.c_uint32),
                ("albedo", C.c_float * 3), ("scale", C.c_float), ("noise_type", C.c_uint32), ("image_idx", C.c_uint32), ("pad", C.c_uint32 * 2)]


class Image(C.Structure):
    _fields_ = [("texel_offset", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("pad", C.c_uint32)]


class Perlin(C.Structure):
    _fields_ = [("perm_x", C.c_int32 * 256), ("perm_y", C.c_int32 * 256), ("perm_z", C.c_int32 * 256),
                ("vec", (C.c_float * 4) * 256)]


class BvhNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("left_first", C.c_uint32), ("bmax", C.c_float * 3), ("count", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00", C.c_float * 3), ("pixel_delta_u", C.c_float * 3),
                ("pixel_delta_v", C.c_float * 3), ("defocus_disk_u", C.c_float * 3), ("defocus_disk_v", C.c_float * 3),
                ("defocus_angle", C.c_float), ("vfov", C.c_float), ("focus_dist", C.c_float), ("look_at", C.c_float * 3)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_spheres", C.c_uint32), ("n_quads", C.c_uint32), ("n_xforms", C.c_uint32), ("n_instances", C.c_uint32),
        ("n_media", C.c_uint32), ("n_materials", C.c_uint32), ("n_textures", C.c_uint32), ("n_perlin", C.c_uint32),
        ("n_prim_refs", C.c_uint32), ("n_node_pairs", C.c_uint32), ("tlas_root", C.c_uint32), ("n_top_level", C.c_uint32),
        ("spheres", C.POINTER(Sphere)), ("quads", C.POINTER(Quad)), ("xforms", C.POINTER(Xform)),
        ("instances", C.POINTER(Instance)), ("media", C.POINTER(Medium)), ("materials", C.POINTER(Material)),
        ("textures", C.POINTER(Texture)), ("perlin", C.POINTER(Perlin)), ("prim_refs", C.POINTER(C.c_uint32)),
        ("nodes", C.POINTER(BvhNode)),
        ("background", C.c_float * 3), ("min_inv_scale", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
        ("camera", Camera),
        ("n_images", C.c_uint32), ("n_image_texels", C.c_uint32), ("images", C.POINTER(Image)), ("image_texels", C.POINTER(C.c_float)),
        ("has_unified_tlas", C.c_uint32), ("tlas_unified_root", C.c_uint32), ("n_inst_leaves", C.c_uint32), ("pad_unified", C.c_uint32),
        ("inst_leaves", C.POINTER(C.c_uint32)),
        ("has_world_tlas", C.c_uint32), ("tlas_world_root", C.c_uint32), ("inst_bounds", C.POINTER(C.c_float)),
    ]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_depth", C.c_int32), ("frames_per_batch", C.c_int32), ("frame_offset", C.c_int32),
                ("frame_stride", C.c_int32), ("flags", C.c_uint32), ("seed", C.c_uint64), ("n_gpus", C.c_int32),
                ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("frames", C.c_uint64), ("launches", C.c_uint64),
                ("gpu_ms_total", C.c_double), ("gpu_ms_extend", C.c_double), ("gpu_ms_shade", C.c_double),
                ("gpu_ms_other", C.c_double), ("gpu_ms_finish", C.c_double), ("gpu_ms_bvh_build", C.c_double), ("box_pair_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("quad_tests", C.c_uint64), ("instance_visits", C.c_uint64), ("gpu_ms_sort", C.c_double),
                ("stack_overflows", C.c_uint64), ("pending_frames", C.c_uint64), ("n_gpus", C.c_uint32), ("instance_split", C.c_uint32),
                ("gpu_ms_extend_inst", C.c_double), ("max_stack_need", C.c_uint32), ("instance_mode", C.c_uint32),
                ("compact_nodes", C.c_uint32), ("node_inflation", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("point", C.c_float * 3), ("t", C.c_float), ("normal", C.c_float * 3), ("material", C.c_int32),
                ("prim", C.c_uint32), ("instance", C.c_int32), ("front_face", C.c_uint32), ("uv16", C.c_uint32)]


# every symbol include/rt2.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "rt2_scene_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(_P)]),
    "rt2_scene_load_string": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(_P)]),
    "rt2_scene_synthetic_spheres": (C.c_int, [C.c_uint32, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "rt2_scene_destroy": (None, [_P]),
    "rt2_scene_get_desc": (C.c_int, [_P, C.POINTER(SceneDesc)]),
    "rt2_scene_set_dims": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "rt2_scene_set_perlin": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "rt2_scene_get_perlin": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "rt2_scene_span1_flags": (C.c_int, [_P, _P]),
    "rt2_create": (C.c_int, [_P, C.POINTER(Config), C.POINTER(_P)]),
    "rt2_destroy": (None, [_P]),
    "rt2_upload_scene": (C.c_int, [_P, _P]),
    "rt2_resize": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "rt2_reset": (C.c_int, [_P]),
    "rt2_update": (C.c_int, [_P, C.c_uint32]),
    "rt2_flush": (C.c_int, [_P]),
    "rt2_synchronize": (C.c_int, [_P]),
    "rt2_frame_idx": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "rt2_dims": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rt2_read_mean_rgb32f": (C.c_int, [_P, _P]),
    "rt2_read_rgba8": (C.c_int, [_P, _P]),
    "rt2_read_accum": (C.c_int, [_P, _P, _P]),
    "rt2_write_accum": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "rt2_accum_device_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rt2_set_frame_idx": (C.c_int, [_P, C.c_uint64]),
    "rt2_accum_ipc_handle": (C.c_int, [_P, _P]),
    "rt2_resolve_peers": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint64, _P, _P]),
    "rt2_intersect": (C.c_int, [_P, _P, C.c_size_t, C.c_float, C.c_float, C.c_int, _P]),
    "rt2_texture_value": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_size_t, _P]),
    "rt2_read_bvh": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rt2_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "rt2_debug_counters": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "rt2_read_queue_sizes": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "rt2_set_profiling": (C.c_int, [_P, C.c_int]),
    "rt2_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "rt2_write_image": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_char_p, C.c_int]),
    "rt2_tonemap_rgb8": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "rt2_app_run": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p]),
    "rt2_last_error": (C.c_char_p, []),
    "rt2_abi_version": (C.c_int, []),
    "rt2_device_count": (C.c_int, []),
    "rt2_measure_fp32_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "rt2_measure_l2_bandwidth": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libraytrace2_b200.so and set prototypes.  Raises if the native library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"native library {LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no Python/CPU fallback for the render path)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class Rt2Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rt2 error {code}: {message}")
        self.code = code
        self.message = message


def check(rc: int) -> None:
    if rc != RT2_OK:
        raise Rt2Error(rc, load_library().rt2_last_error().decode("utf-8", "replace"))
