"""CPU: the round-2 pins (tests/golden/make_golden_r2.py).  The legacy-format scenes cannot be loaded by the reference at HEAD;
tools/convert_legacy.py rewrites them into the current format, the UNMODIFIED reference (oracle/_ref) produced hits / moments /
texture values from the converted files, and here our CPU restatement — loading the ORIGINAL legacy files through its own
adapter — must reproduce them bit for bit.  Where oracle/_ref exists (the build container) the converter is also checked live."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, scene_path

sys.path.insert(0, os.path.join(ROOT, "tools"))
import convert_legacy  # noqa: E402

LEGACY_GOLDEN = ["final_render_book_1", "light_scene1", "checker_test", "cornell_box2"]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", LEGACY_GOLDEN)
def test_port_on_legacy_file_equals_reference_on_converted_file(port_oracle, name):
    g = np.load(os.path.join(GOLDEN, f"hits_{name}.npz"))
    port = port_oracle.PortScene(scene_path(name), 16)
    r = port.intersect(g["origins"], g["directions"], g["times"])
    hit = g["hit"].astype(bool)
    assert hit.sum() > 500 and np.array_equal(r["hit"].astype(bool), hit)
    for key in ("t", "point", "normal"):
        assert np.array_equal(_bits(r[key][hit]), _bits(g[key][hit])), key
    assert np.array_equal(r["material"][hit], g["material"][hit])
    assert np.array_equal(r["front_face"][hit], g["front_face"][hit])


@pytest.mark.parametrize("name", ["final_render_book_1", "light_scene1", "cornell_box2"])
def test_port_render_of_legacy_scene_matches_reference_moments(port_oracle, name):
    from raytrace2_b200 import parity
    g = np.load(os.path.join(GOLDEN, f"moments_{name}.npz"))
    dims, spp = tuple(int(x) for x in g["dims"]), int(g["spp"])
    port = port_oracle.PortScene(scene_path(name), spp, dims=dims)
    if "perm_x" in g.files:
        port.perlin_set(int(g["noise_tex"]), g["perm_x"], g["perm_y"], g["perm_z"], g["vec"])
    s, ss, rays, _ = port.render(0, spp, 50, 0, True)
    z, valid = parity.z_scores(s, ss, spp, g["sum"].astype(np.float64), g["sumsq"].astype(np.float64), spp)
    st = parity.summary(z, valid)
    assert st["n"] > 1000 and abs(st["mean_z"]) < 4.0 / np.sqrt(st["n"]) + 0.01, st
    assert 0.90 < st["std_z"] < 1.06 and st["frac_gt3"] < 0.005, st
    assert abs(rays - float(g["rays"])) < 0.01 * float(g["rays"])


def test_port_textures_equal_reference_at_fixed_points(port_oracle, tmp_path):
    for tag in ("checker", "noise"):
        g = np.load(os.path.join(GOLDEN, f"texture_{tag}.npz"))
        path = tmp_path / f"{tag}.json"
        path.write_bytes(g["scene_json"].tobytes())
        port = port_oracle.PortScene(str(path), 16, data_dir=os.path.join(ROOT, "data"))
        for key in [k for k in g.files if k.startswith("value_")]:
            ti = int(key.split("_")[1])
            if tag == "noise":
                port.perlin_set(ti, g[f"perm_x_{ti}"], g[f"perm_y_{ti}"], g[f"perm_z_{ti}"], g[f"vec_{ti}"])
            assert np.array_equal(_bits(port.texture_value(ti, g["points"])), _bits(g[key])), (tag, ti)


def test_converter_output_is_current_format():
    for name in LEGACY_GOLDEN + ["scene2", "quad_scene1", "final_render_scene_blur", "perlin_spheres"]:
        doc = json.load(open(scene_path(name)))
        assert convert_legacy.is_legacy(doc)
        out = convert_legacy.convert(doc, name + ".json", os.path.join(ROOT, "data"))
        assert isinstance(out["primitives"], list) and len(out["scene"]) == len(out["primitives"])
        assert all("id" not in m and m["type"] for m in out["materials"])
        assert all(0 <= p["material"] < len(out["materials"]) for p in out["primitives"])
        assert isinstance(out["camera"], (dict, str))
        n_legacy = sum(len(v) for v in doc["primitives"].values())
        assert len(out["primitives"]) == n_legacy
    cur = json.load(open(scene_path("cornell_original_test")))
    assert convert_legacy.convert(cur, "x.json", ".") is cur


def test_reference_loads_converted_legacy_files(ref_oracle, port_oracle, tmp_path):
    """Live in the build container: the unmodified reference on the converted file == our restatement on the legacy file."""
    rng = np.random.default_rng(8)
    for name in ["scene2", "quad_scene1", "final_render_scene_blur"]:
        dst = str(tmp_path / (name + ".json"))
        convert_legacy.convert_file(scene_path(name), dst, os.path.join(ROOT, "data"))
        ref = ref_oracle.RefScene(dst, 16)
        port = port_oracle.PortScene(scene_path(name), 16)
        assert (ref.width, ref.height, ref.n_top) == (port.width, port.height, port.n_top)
        o = rng.uniform(-12, 12, size=(20000, 3)).astype(np.float32)
        v = rng.normal(size=(20000, 3))
        d = (v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0.05, 2.0, (20000, 1))).astype(np.float32)
        t = rng.random(20000).astype(np.float32)
        a, b = ref.intersect(o, d, t), port.intersect(o, d, t)
        assert np.array_equal(a["hit"], b["hit"]) and a["hit"].sum() > 300
        h = a["hit"].astype(bool)
        for key in ("t", "point", "normal"):
            assert np.array_equal(_bits(a[key][h]), _bits(b[key][h])), (name, key)
        assert np.array_equal(a["material"][h], b["material"][h])
