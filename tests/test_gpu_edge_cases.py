"""GPU: degenerate inputs of the hot path — an empty scene, a scene that holds only a constant medium, a single primitive,
odd image sizes with batches that do not divide the sample count, max_depth 1, and both extend kernels on the same input."""
import os

import numpy as np
import pytest

import raytrace2_b200 as rt
from raytrace2_b200 import scene_builder as sb

pytestmark = pytest.mark.gpu


def test_empty_scene_renders_the_background(native_lib):
    b = sb.SceneBuilder(width=37, aspect_ratio=37 / 23, background=(0.2, 0.3, 0.4))
    tr = rt.RayTracer(rt.Scene.from_builder(b), num_samples=9, seed=1)
    assert tr.Dims() == (37, 23)
    tr.Update(9)
    img = tr.NonConvertedPixels()
    assert img.shape == (23, 37, 3)
    assert np.allclose(img, np.float32([0.2, 0.3, 0.4]), rtol=1e-6)
    st = tr.stats()
    assert st["paths"] == 37 * 23 * 9 and st["rays"] == st["paths"]  # every camera ray misses, nothing scatters
    hits = tr.intersect(np.zeros((5, 3), np.float32), np.tile(np.float32([0, 0, -1]), (5, 1)))
    assert np.all(hits["material"] == -1) and np.all(hits["prim"] == 0xFFFFFFFF)


def test_medium_only_scene(native_lib):
    """No surface at all: the tree is empty and every hit comes from ConstantMedium::Hit."""
    b = sb.SceneBuilder(width=48, fov=40, center=(0, 0, 6), look_at=(0, 0, 0), background=(1, 1, 1))
    b.place(b.sphere((0, 0, 0), 1.5, b.lambertian((0.5, 0.5, 0.5)), medium=b.constant_medium(2.0, (0.9, 0.1, 0.1))))
    tr = rt.RayTracer(rt.Scene.from_builder(b), num_samples=64, seed=3)
    tr.Update(64)
    img = tr.NonConvertedPixels()
    centre, corner = img[24, 24], img[1, 1]
    assert np.allclose(corner, 1.0, atol=1e-6)                       # rays that miss the ball see the background
    assert centre[0] > centre[1] + 0.1 and centre[0] < 1.0          # the red medium tints (and darkens) the centre
    assert tr.stats()["rays"] > tr.stats()["paths"]


@pytest.mark.parametrize("flags", [0, rt.RT2_FLAG_NO_FLAT_EXTEND])
def test_single_primitive_and_depth_one(native_lib, flags):
    b = sb.SceneBuilder(width=64, fov=40, center=(0, 0, 5), look_at=(0, 0, 0), background=(0.5, 0.7, 1.0))
    b.place(b.sphere((0, 0, 0), 1.0, b.lambertian((0.8, 0.8, 0.8))))
    scene = rt.Scene.from_builder(b)
    deep = rt.RayTracer(scene, num_samples=16, seed=5, max_depth=50, flags=flags)
    one = rt.RayTracer(scene, num_samples=16, seed=5, max_depth=1, flags=flags)
    deep.Update(16)
    one.Update(16)
    a, c = deep.NonConvertedPixels(), one.NonConvertedPixels()
    # RayColor(depth <= 0) is black (RayTracer.cpp:21-23): with max_depth 1 the sphere is black, the sky unchanged
    assert np.allclose(c[32, 32], 0.0) and np.allclose(c[2, 2], a[2, 2])
    assert a[32, 32].min() > 0.1
    assert one.stats()["rays"] == one.stats()["paths"]
    g = deep.intersect(np.float32([[0, 0, 5], [0, 3, 5]]), np.float32([[0, 0, -1], [0, 0, -1]]))
    assert g["material"][0] == 0 and g["t"][0] == np.float32(4.0) and g["material"][1] == -1


def test_odd_sizes_and_ragged_batches(native_lib):
    """13 frames in batches of 5 + 5 + 3 on a 61 x 19 image equal one batch of 13, bit for bit."""
    b = sb.cornell_box(width=61)
    scene = rt.Scene.from_builder(b)
    kw = dict(num_samples=13, seed=9, dims=(61, 19))
    a = rt.RayTracer(scene, frames_per_batch=5, **kw)
    c = rt.RayTracer(scene, frames_per_batch=13, **kw)
    a.Update(13)
    for _ in range(13):
        c.Update(1)
    assert a.FrameIdx() == c.FrameIdx() == 13
    assert np.array_equal(a.read_accum().view(np.uint32), c.read_accum().view(np.uint32))
    assert a.stats()["rays"] == c.stats()["rays"]
