"""The C-ABI library loads and exports every symbol include/rt2.h declares; struct layouts match (no compute calls)."""
import ctypes as C
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rt2.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(native_lib):
    from raytrace2_b200 import _capi
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in include/rt2.h but not exported"
        assert s in _capi.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_capi.PROTOTYPES) == set(syms)


def test_abi_version_and_error_string(native_lib):
    assert native_lib.rt2_abi_version() == 3
    assert isinstance(native_lib.rt2_last_error(), bytes)


def test_struct_sizes():
    from raytrace2_b200 import _capi
    assert C.sizeof(_capi.Sphere) == 32
    assert C.sizeof(_capi.Quad) == 80
    assert C.sizeof(_capi.Xform) == 96
    assert C.sizeof(_capi.Instance) == 16
    assert C.sizeof(_capi.Medium) == 32
    assert C.sizeof(_capi.Material) == 32
    assert C.sizeof(_capi.Texture) == 48
    assert C.sizeof(_capi.Perlin) == 3 * 1024 + 4096
    assert C.sizeof(_capi.BvhNode) == 32
    assert C.sizeof(_capi.Hit) == 48


def test_errors_do_not_throw(native_lib):
    from raytrace2_b200 import _capi
    h = C.c_void_p()
    rc = native_lib.rt2_scene_load(b"/nonexistent/scene.json", None, 0, C.byref(h))
    assert rc == _capi.RT2_ERR_IO and not h.value
    assert b"Failed to open json file" in native_lib.rt2_last_error()
    rc = native_lib.rt2_scene_load_string(b"{ not json", None, 0, C.byref(h))
    assert rc == _capi.RT2_ERR_PARSE
    rc = native_lib.rt2_create(None, None, C.byref(h))
    assert rc == _capi.RT2_ERR_INVALID_ARG


def test_no_cpu_fallback_without_device(native_lib):
    """On a box without a GPU, creating a renderer must fail loudly with RT2_ERR_CUDA (never render on the CPU)."""
    import raytrace2_b200 as rt
    if native_lib.rt2_device_count() > 0:
        return
    scene = rt.Scene.load(os.path.join(ROOT, "data", "cornell_original_test.json"))
    try:
        rt.RayTracer(scene)
    except rt.Rt2Error as e:
        assert e.code == -4
    else:
        raise AssertionError("renderer was created without a CUDA device")


def test_product_does_not_import_oracle():
    """The product tree must never reference oracle/ (parity claims are void otherwise)."""
    pkg = os.path.join(ROOT, "raytrace2_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f in ("parity.py",), f"{f} mentions the oracle"
