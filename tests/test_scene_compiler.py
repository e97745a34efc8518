"""Host scene compiler (C++ behind rt2_scene_*): loader defaults and quirks, legacy adapter, flattening, BVH invariants,
camera block and Q2 flags against the golden vectors written by the real reference, image writer."""
import json
import os
import struct
import zlib

import numpy as np
import pytest

import raytrace2_b200 as rt
from conftest import CURRENT_SCENES, GOLDEN, LEGACY_SCENES, scene_path
from raytrace2_b200 import _capi


@pytest.mark.parametrize("name", CURRENT_SCENES)
def test_camera_block_and_span1_match_reference_golden(native_lib, name):
    scene = rt.Scene.load(scene_path(name))
    d = scene.desc
    cam = np.load(os.path.join(GOLDEN, f"camera_{name}.npy"))
    mine = np.concatenate([np.array(list(getattr(d.camera, k)), np.float32) for k in
                           ["center", "pixel00", "pixel_delta_u", "pixel_delta_v", "defocus_disk_u", "defocus_disk_v"]])
    assert np.array_equal(mine.view(np.uint32), cam[:18].view(np.uint32)), "camera block must be bit-identical with Camera::Update"
    span1 = np.load(os.path.join(GOLDEN, f"span1_{name}.npy"))
    assert np.array_equal(scene.span1_flags(), span1)


def test_book2_counts_and_q2_fog(native_lib):
    """409 top-level objects = 400 boxes (2400 quads) + light quad + 7 spheres + instance of 1000 spheres; the
    r=5000 fog (top-level 406) is in a span-1 leaf => sampled twice; the blue medium (405) is not."""
    scene = rt.Scene.load(scene_path("book2_final_scene_10000_samples"))
    d = scene.desc
    assert (d.n_top_level, d.n_quads, d.n_spheres, d.n_instances, d.n_media) == (409, 2401, 1007, 1, 2)
    assert (d.n_materials, d.n_textures, d.n_perlin) == (9, 4, 1)
    media = scene.media()
    by_top = {int(m["top_level_node"]): m for m in media}
    assert by_top[406]["sample_twice"] == 1 and by_top[405]["sample_twice"] == 0
    assert by_top[406]["neg_inv_density"] == np.float32(-1.0 / np.float32(1e-4))
    assert scene.span1_flags().sum() == 103
    assert (d.width, d.height) == (600, 600)


def test_cornell_volume_media_under_instances(native_lib):
    scene = rt.Scene.load(scene_path("cornell_volume_10000_samples"))
    d = scene.desc
    assert d.n_media == 2 and d.n_instances == 0 and d.n_quads == 6 + 12
    for m in scene.media():
        assert m["chain_len"] == 1 and m["boundary_count"] == 6 and m["sample_twice"] == 0
    mats = scene.materials()
    assert [int(m["type"]) for m in mats[-2:]] == [_capi.MAT_ISOTROPIC] * 2  # appended implicitly (Serialize.cpp:324-331)


def test_nested_transforms_flatten_to_chains(native_lib):
    scene = rt.Scene.load(scene_path("cornell_box_scene_graph"))
    inst = scene.instances()
    assert sorted(int(i["chain_len"]) for i in inst) == [1, 2, 3]
    assert scene.desc.n_xforms == 6  # chains stored contiguously: 1 + 2 + 3 levels
    x = scene.xforms()
    # a translation-only level (no "rotation" key) must be an exact identity rotation
    lvl = x[int([i for i in inst if i["chain_len"] == 3][0]["chain_first"]) + 2]
    m = np.array(lvl["model"])
    assert np.array_equal(m[:, :3], np.eye(3, dtype=np.float32))


def _check_bvh(scene):
    d = scene.desc
    nodes, refs = scene.nodes(), scene.prim_refs()
    seen = np.zeros(len(refs), np.int32)
    media_refs = set()
    for m in scene.media():
        media_refs.update(range(int(m["boundary_first"]), int(m["boundary_first"] + m["boundary_count"])))

    def walk(pair, lo, hi, depth):
        assert depth <= 64
        for side in range(2):
            n = nodes[2 * pair + side]
            bmin, bmax = np.array(n["bmin"]), np.array(n["bmax"])
            if not (bmin[0] <= bmax[0]):
                continue  # empty slot (NaN bounds)
            if lo is not None:
                assert np.all(bmin >= lo - 1e-3) and np.all(bmax <= hi + 1e-3), "child box must lie inside its parent"
            if n["count"] == 0:
                walk(int(n["left_first"]), bmin, bmax, depth + 1)
            else:
                assert 1 <= n["count"] <= 16
                ids = range(int(n["left_first"]), int(n["left_first"] + n["count"]))
                types = [int(refs[i]) >> 28 for i in ids]
                if _capi.RT2_PRIM_INSTANCE in types:
                    assert n["count"] == 1, "instance references must be singleton leaves"
                for i in ids:
                    seen[i] += 1
    roots = [int(d.tlas_root)] + [int(i["blas_root"]) for i in scene.instances()]
    for r in roots:
        walk(r, None, None, 0)
    n_inst = len(scene.instances())
    tlas_slots = _leaf_ref_slots(nodes, int(d.tlas_root))
    surfaces = sorted(int(refs[i]) for i in tlas_slots if (int(refs[i]) >> 28) != _capi.RT2_PRIM_INSTANCE)
    extra = np.zeros(len(refs), np.int32)
    if d.has_world_tlas:
        # instance split: a second world tree over the same surfaces, without the instance leaves
        assert 1 <= n_inst <= _capi.RT2_MAX_HOISTED_INSTANCES
        before = seen.copy()
        walk(int(d.tlas_world_root), None, None, 0)
        world_only = seen - before
        seen[:] = before
        extra += world_only
        assert sorted(int(refs[i]) for i in range(len(refs)) if world_only[i]) == surfaces, \
            "the surfaces-only world tree must hold exactly the TLAS's non-instance leaves"
        b = np.ctypeslib.as_array(d.inst_bounds, shape=(n_inst, 8))
        assert np.all(b[:, 0:3] < b[:, 4:7])
    else:
        assert n_inst == 0 or n_inst > _capi.RT2_MAX_HOISTED_INSTANCES
    if d.has_unified_tlas:
        # unified world tree: the surfaces + every instanced primitive once, as leaf (INSTANCE << 28 | k) -> inst_leaves[k]
        before = seen.copy()
        walk(int(d.tlas_unified_root), None, None, 0)
        uni = seen - before
        seen[:] = before
        extra += uni
        got = sorted(int(refs[i]) for i in range(len(refs)) if uni[i])
        inst_refs = [r for r in got if (r >> 28) == _capi.RT2_PRIM_INSTANCE]
        assert [r for r in got if (r >> 28) != _capi.RT2_PRIM_INSTANCE] == surfaces
        assert [r & 0x0FFFFFFF for r in inst_refs] == list(range(d.n_inst_leaves)) and d.n_inst_leaves > 0
        leaves = np.ctypeslib.as_array(d.inst_leaves, shape=(d.n_inst_leaves, 2))
        assert np.all(leaves[:, 1] < n_inst) and np.all((leaves[:, 0] >> 28) <= _capi.RT2_PRIM_QUAD)
        # every BLAS primitive of every instance appears exactly once
        want = []
        for j, inst in enumerate(scene.instances()):
            want += [(int(refs[i]), j) for i in _leaf_ref_slots(nodes, int(inst["blas_root"]))]
        assert sorted(want) == sorted((int(a), int(b)) for a, b in leaves)
    else:
        assert n_inst == 0
    for i in range(len(refs)):
        assert seen[i] + extra[i] == (0 if i in media_refs else 1), f"prim ref {i} referenced {seen[i] + extra[i]} times"


def _leaf_ref_slots(nodes, root):
    """prim_refs slots reachable from one tree."""
    out, todo = set(), [root]
    while todo:
        pair = todo.pop()
        for side in range(2):
            n = nodes[2 * pair + side]
            if not (n["bmin"][0] <= n["bmax"][0]):
                continue
            if n["count"] == 0:
                todo.append(int(n["left_first"]))
            else:
                out.update(range(int(n["left_first"]), int(n["left_first"] + n["count"])))
    return out


@pytest.mark.parametrize("name", CURRENT_SCENES + ["final_render_book_1"])
def test_bvh_invariants(native_lib, name):
    _check_bvh(rt.Scene.load(scene_path(name)))


def test_quad_constants_follow_reference_constructor(native_lib):
    """Quad ctor (Quad.hpp:14-21): n = cross(u,v); normal = normalize(n); d = dot(normal,q); w = n / dot(n,n) in float32."""
    scene = rt.Scene.load(scene_path("cornell_original_test"))
    f = np.float32
    for q in scene.quads():
        u, v, qq = (np.array(q[k], f) for k in ("u", "v", "q"))
        n = np.array([u[1] * v[2] - v[1] * u[2], u[2] * v[0] - v[2] * u[0], u[0] * v[1] - v[0] * u[1]], f)
        nn = f(f(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
        normal = n * f(f(1) / np.sqrt(nn))
        dd = f(f(normal[0] * qq[0] + normal[1] * qq[1]) + normal[2] * qq[2])
        assert np.array_equal(np.array(q["normal"], f).view(np.uint32), normal.view(np.uint32))
        assert f(q["d"]) == dd
        assert np.array_equal(np.array(q["w"], f), n / nn)


@pytest.mark.parametrize("name", LEGACY_SCENES)
def test_legacy_adapter(native_lib, name):
    """13 scene files of the reference are in a legacy format HEAD's loader throws on; the adapter turns every
    primitive into a top-level node and picks the documented camera."""
    scene = rt.Scene.load(scene_path(name))
    doc = json.load(open(scene_path(name)))
    pr = doc["primitives"]
    d = scene.desc
    n_sph, n_quad, n_box = len(pr.get("spheres", [])), len(pr.get("quads", [])), len(pr.get("boxes", []))
    assert d.n_spheres == n_sph and d.n_quads == n_quad + 6 * n_box
    assert d.n_top_level == n_sph + n_quad + n_box
    if name.startswith("final_render"):
        assert list(d.camera.center) == [13.0, 2.0, 3.0] and d.camera.vfov == 20.0  # data/cam1.json
        assert (d.width, d.height) == (1600, 900)  # App.cpp:115


def test_loader_defaults_and_quirks(native_lib):
    doc = {"camera": {"fov": 40.9, "width": 300, "aspect_ratio": 1.5},
           "materials": [{"type": "lambertian"}, {"type": "diffuse_light", "albedo": [4, 4, 4]}, {"type": "texture", "albedo": [0.1, 0.2, 0.3]},
                         {"type": "metal"}, {"type": "dielectric"}],
           "primitives": [{"type": "sphere", "material": 0}, {"type": "bogus"}, {"type": "quad", "material": 1},
                          {"type": "box", "material": 3, "constant_medium": {"albedo": [1, 0, 0]}}],
           "scene": [{"primitive": 0}, {"primitive": 1}, {"primitive": 2}]}
    scene = rt.Scene.from_string(json.dumps(doc))
    d = scene.desc
    assert d.camera.vfov == 40.0, "int-typed default truncates a fractional fov (Serialize.cpp:34)"
    assert (d.width, d.height) == (300, 200)
    assert list(d.background) == [1.0, 1.0, 1.0]
    assert list(d.camera.center) == [0.0, 0.0, 1.0] and d.camera.focus_dist == 1.0 and d.camera.defocus_angle == 0.0
    sph = scene.spheres()[0]
    assert sph["radius"] == 0.5 and list(sph["center0"]) == [0, 0, 0]
    # the invalid primitive is skipped, so scene index 1 is the quad and 2 the (medium) box
    assert d.n_quads == 1 + 6 and d.n_media == 1
    q = scene.quads()[0]
    assert list(q["u"]) == [1, 0, 0] and list(q["v"]) == [0, 0, 1] and q["material"] == 1
    mats, texs = scene.materials(), scene.textures()
    assert len(mats) == 6 and int(mats[5]["type"]) == _capi.MAT_ISOTROPIC  # appended by the constant_medium block
    assert len(texs) == 3  # light albedo, texture albedo, medium albedo -> implicit SolidColor textures
    assert list(texs[int(mats[1]["tex_idx"])]["albedo"]) == [4, 4, 4]
    assert scene.media()[0]["neg_inv_density"] == np.float32(-1.0 / np.float32(0.01))
    assert mats[4]["refraction_index"] == 1.0 and mats[3]["fuzz"] == 0.0


def test_loader_errors(native_lib):
    with pytest.raises(rt.Rt2Error) as e:
        rt.Scene.from_string(json.dumps({"camera": {}, "materials": [{"albedo": [1, 1, 1]}], "primitives": [], "scene": []}))
    assert e.value.code == _capi.RT2_ERR_PARSE and "material type field empty" in e.value.message
    with pytest.raises(rt.Rt2Error):
        rt.Scene.from_string(json.dumps({"camera": {}, "materials": [], "primitives": [], "scene": [{"primitive": 3}]}))
    with pytest.raises(rt.Rt2Error):
        rt.Scene.from_string(json.dumps({"camera": {}, "materials": [], "primitives": [{"type": "sphere", "material": 2}], "scene": [{"primitive": 0}]}))
    assert rt.SceneLoader().LoadScene("/nonexistent.json") is None
    # empty scene: loads, nothing to hit
    s = rt.Scene.from_string(json.dumps({"camera": {}, "materials": [], "primitives": [], "scene": []}))
    assert s.desc.n_top_level == 0 and s.desc.n_node_pairs == 1


def test_named_camera_file(native_lib, tmp_path):
    (tmp_path / "mycam.json").write_text(json.dumps({"fov": 33, "center": [1, 2, 3], "look_at": [0, 1, 0], "focus_distance": 2.5}))
    doc = {"camera": "mycam", "materials": [{"type": "lambertian"}], "primitives": [{"type": "sphere", "material": 0}], "scene": [{"primitive": 0}]}
    (tmp_path / "s.json").write_text(json.dumps(doc))
    d = rt.Scene.load(str(tmp_path / "s.json")).desc
    assert d.camera.vfov == 33.0 and list(d.camera.center) == [1, 2, 3] and d.camera.focus_dist == 2.5
    assert (d.width, d.height) == (1600, 900)


def test_synthetic_sphere_scene(native_lib):
    s = rt.Scene.synthetic_spheres(5000, seed=7, width=320, height=180)
    d = s.desc
    assert d.n_spheres == 5001 and d.n_materials == 5001 and d.n_instances == 0 and d.n_media == 0
    sph = s.spheres()
    assert sph[0]["radius"] == 100000.0
    c = np.array([x["center0"] for x in sph[1:]])
    assert c[:, 0].min() >= -1000 and c[:, 0].max() <= 1000 and c[:, 1].min() >= 0 and c[:, 1].max() <= 200
    types = np.array([m["type"] for m in s.materials()[1:]])
    assert 0.7 < (types == 0).mean() < 0.9 and 0.1 < (types == 1).mean() < 0.2
    s2 = rt.Scene.synthetic_spheres(5000, seed=7, width=320, height=180)
    assert np.array_equal(s2.spheres(), sph), "generator must be deterministic in its seed"
    _check_bvh(s)


def _decode_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat = 8, b""
    while pos < len(data):
        ln, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        crc = struct.unpack(">I", data[pos + 8 + ln:pos + 12 + ln])[0]
        assert crc == zlib.crc32(tag + body) & 0xFFFFFFFF
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + ln
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, w * 3 + 1)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 3)


def test_write_image_matches_reference_golden(native_lib, tmp_path):
    """util::WriteImage (Util.cpp:39-79): sqrt gamma, clamp(255.999*c, 0, 255), vertical flip — golden bytes decoded
    from the PNG the real reference wrote for the same float image."""
    g = np.load(os.path.join(GOLDEN, "tonemap.npz"))
    img, want = g["image"], g["rgb8"]
    from raytrace2_b200.raytracer import tonemap_rgb8
    with np.errstate(invalid="ignore"):
        got = tonemap_rgb8(img)
    assert np.array_equal(got, want)
    out = tmp_path / "o.png"
    rt.WriteImage(img, img.shape[1], img.shape[0], str(out), png=True)
    assert np.array_equal(_decode_png(str(out)), want)
    ppm = tmp_path / "o.ppm"
    rt.WriteImage(img, img.shape[1], img.shape[0], str(ppm), png=False)
    toks = open(ppm).read().split()
    assert toks[:4] == ["P3", str(img.shape[1]), str(img.shape[0]), "255"]
    assert np.array_equal(np.array(toks[4:], np.int32).reshape(want.shape), want)


def test_perlin_tables_roundtrip_and_shape(native_lib):
    scene = rt.Scene.load(scene_path("book2_final_scene_10000_samples"), perlin_seed=3)
    px, py, pz, vec = scene.get_perlin(0)
    for p in (px, py, pz):
        assert sorted(p.tolist()) == list(range(256)), "permutation tables (PerlinNoiseGen.cpp:90-103)"
    assert np.allclose(np.linalg.norm(vec, axis=1), 1.0, atol=1e-6), "gradients are normalised (PerlinNoiseGen.cpp:44)"
    other = rt.Scene.load(scene_path("book2_final_scene_10000_samples"), perlin_seed=4)
    assert not np.array_equal(other.get_perlin(0)[0], px)
    other.set_perlin(0, px, py, pz, vec)
    for a, b in zip(other.get_perlin(0), (px, py, pz, vec)):
        assert np.array_equal(a, b)
