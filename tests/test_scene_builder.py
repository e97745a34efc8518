"""Scene authoring (raytrace2_b200/scene_builder.py): the generated documents load through the same host
scene compiler as the repo's files, and the generated Cornell box flattens to the same geometry as data/cornell_original_test.json."""
import json

import numpy as np

import raytrace2_b200 as rt
from conftest import scene_path
from raytrace2_b200 import scene_builder as sb


def _quad_set(scene):
    q = scene.quads()
    rows = np.concatenate([np.asarray(q["q"]), np.asarray(q["u"]), np.asarray(q["v"])], axis=1)
    return sorted(map(tuple, np.round(rows, 4).tolist()))


def test_cornell_box_builder_matches_repo_scene(native_lib):
    mine = rt.Scene.from_builder(sb.cornell_box())
    repo = rt.Scene.load(scene_path("cornell_original_test"))
    dm, dr = mine.desc, repo.desc
    assert (dm.n_quads, dm.n_instances, dm.n_spheres, dm.n_media) == (dr.n_quads, dr.n_instances, dr.n_spheres, dr.n_media) == (18, 2, 0, 0)
    assert (dm.width, dm.height) == (dr.width, dr.height) == (600, 600)
    assert _quad_set(mine) == _quad_set(repo)
    # the camera block is computed from the same parameters
    for f in ("center", "pixel00", "pixel_delta_u", "pixel_delta_v"):
        assert np.array_equal(np.asarray(getattr(dm.camera, f)), np.asarray(getattr(dr.camera, f))), f


def test_cornell_volume_and_book2_builders(native_lib):
    vol = rt.Scene.from_builder(sb.cornell_volume())
    repo = rt.Scene.load(scene_path("cornell_volume_10000_samples"))
    assert vol.desc.n_media == repo.desc.n_media == 2 and vol.desc.n_quads == repo.desc.n_quads
    assert _quad_set(vol) == _quad_set(repo)
    b2 = rt.Scene.from_builder(sb.book2_final(seed=3))
    d = b2.desc
    assert d.n_quads == 400 * 6 + 1 and d.n_spheres == 7 + 1000 and d.n_media == 2 and d.n_instances == 1 and d.n_perlin == 1
    assert d.n_top_level == 400 + 1 + 7 + 1
    # same seed, same document; another seed, another box grid
    assert sb.book2_final(seed=3).to_json() == sb.book2_final(seed=3).to_json()
    assert sb.book2_final(seed=4).to_json() != sb.book2_final(seed=3).to_json()


def test_random_spheres_and_builder_primitives(native_lib):
    s = rt.Scene.from_builder(sb.random_spheres(100, width=320))
    assert s.desc.n_spheres == 100 + 4 and (s.desc.width, s.desc.height) == (320, 180)
    b = sb.SceneBuilder(width=64, aspect_ratio=2.0, background=(0.1, 0.2, 0.3))
    tex = b.checker(0.5, b.solid((1, 1, 1)), b.solid((0, 0, 0)))
    inner = b.node(b.sphere((0, 0, 0), 1.0, b.textured(tex)), transform=b.transform(scale=(2, 2, 2)))
    b.place(transform=b.transform(translation=(1, 2, 3), rotation_deg_axis=(30, 0, 1, 0)), children=[inner])
    b.place(b.quad((0, 0, 0), (1, 0, 0), (0, 1, 0), b.light(tex_idx=tex)))
    doc = json.loads(b.to_json())
    assert set(doc) == {"textures", "materials", "primitives", "scene", "camera", "background_color"}
    sc = rt.Scene.from_builder(b)
    d = sc.desc
    assert d.n_instances == 1 and d.n_xforms == 2 and d.n_textures >= 3 and (d.width, d.height) == (64, 32)
    assert np.allclose(np.asarray(d.background), [0.1, 0.2, 0.3])


def test_write_settings(tmp_path):
    p = tmp_path / "settings.json"
    sb.write_settings(str(p), num_samples=64, max_depth=12)
    assert json.loads(p.read_text()) == {"num_samples": 64, "max_depth": 12, "render_once": True, "save_after_render_once": True,
                                         "render_window": False}
