"""GPU: EVERY scene file the reference ships in data/ (22 files: 9 in the current format, 13 in the legacy format) renders
through the C ABI and agrees with the CPU restatement's render of the same file at equal spp — per-pixel z-scores,
rays per path — at a size the restatement finishes in about a second.  The BASELINE configs have larger dedicated
tests; this sweep is about coverage of the loader paths, camera files, checker / Perlin textures, emitters, motion blur,
touching spheres, quads and volumes as the reference's own scenes combine them."""
import glob
import json
import os

import numpy as np
import pytest

import raytrace2_b200 as rt
from conftest import DATA
from raytrace2_b200 import parity

pytestmark = pytest.mark.gpu


def _scene_files():
    out = []
    for f in sorted(glob.glob(os.path.join(DATA, "*.json"))):
        with open(f) as fh:
            doc = json.load(fh)
        if "materials" in doc:  # camera files have no scene content
            out.append(os.path.basename(f)[:-5])
    return out


SCENE_FILES = _scene_files()
# data/final_render_checker.json gives its ground material a `tex_idx` but carries no "textures" array: the reference
# would index an empty vector (Material.cpp:64-66, undefined behaviour); the loader rejects the file instead
BROKEN = {"final_render_checker": "texture index out of range"}


def test_scene_inventory():
    assert len(SCENE_FILES) == 22


@pytest.mark.parametrize("name", SCENE_FILES)
def test_every_reference_scene_matches_the_restatement(native_lib, port_oracle, name):
    path = os.path.join(DATA, name + ".json")
    if name in BROKEN:
        with pytest.raises(rt.Rt2Error) as ei:
            rt.Scene.load(path)
        assert ei.value.code == -3 and BROKEN[name] in ei.value.message
        return
    scene = rt.Scene.load(path)
    w, h = scene.dims
    dims = (96, max(16, int(round(96 * h / w))))
    spp = 100
    port = port_oracle.PortScene(path, spp, dims=dims)
    if scene.desc.n_perlin:
        for k, ti in enumerate(port.noise_textures):
            port.perlin_set(ti, *scene.get_perlin(min(k, scene.desc.n_perlin - 1)))
    tracer = rt.RayTracer(scene, num_samples=spp, max_depth=50, seed=31, flags=rt.RT2_FLAG_MOMENTS, dims=dims)
    tracer.Update(spp)
    s, ss = tracer.read_accum(moments=True)
    rs, rss, rrays, _ = port.render(0, spp, 50, 0, True)
    assert np.all(np.isfinite(s))
    paths = dims[0] * dims[1] * spp
    st = tracer.stats()
    assert st["paths"] == paths
    rpp_gpu, rpp_ref = st["rays"] / paths, rrays / paths
    assert abs(rpp_gpu - rpp_ref) < 0.02 * rpp_ref + 0.02, (name, rpp_gpu, rpp_ref)
    z, valid = parity.z_scores(s, ss, spp, rs, rss, spp)
    zs = parity.summary(z, valid)
    if zs["n"] > 500:
        assert abs(zs["mean_z"]) < 5.0 / np.sqrt(zs["n"]) + 0.02, (name, zs)
        assert zs["std_z"] < 1.08 and zs["frac_gt3"] < 0.008 and zs["frac_gt4"] < 0.0015, (name, zs)
    # pixels without usable variance (constant background, unlit black): the means must simply agree
    flat = ~valid
    if flat.any():
        assert np.allclose((s / spp)[flat], (rs / spp)[flat], rtol=1e-3, atol=2e-3), name
