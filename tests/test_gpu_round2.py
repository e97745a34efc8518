"""GPU tests of the round-2 work, all through the C ABI:

  * the two-pass instance split (k_traverse<kTravWorld> + k_traverse<kTravInst>) against the inline walk: same closest hits;
  * rt2_update called once per sample (the reference's own loop, App.cpp:243-248) = one big call, bit for bit, batched;
  * one handle over several GPUs (rt2_config.n_gpus): same image as one GPU up to fp32 summation order;
  * device texture code (rt2_texture_value) against the reference's Texture::Value at fixed points;
  * the reference's exact-tie rule (later quad wins, HittableList.cpp:8-22);
  * legacy scenes against goldens rendered by the REAL reference from converted files (tools/convert_legacy.py);
  * BASELINE config 1 at full size (600 x 600, 1024 spp) against the reference's tile moments;
  * the traversal-stack overflow counter."""
import json
import os

import numpy as np
import pytest

import raytrace2_b200 as rt
from _rays import fixed_rays
from conftest import GOLDEN, scene_path
from raytrace2_b200 import parity

pytestmark = pytest.mark.gpu

BOOK2 = "book2_final_scene_10000_samples"


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- instances: unified world tree (default), two-pass split, inline TLAS -> BLAS ----------------------------------------
UNIFIED, SPLIT, INLINE = 0, rt.RT2_FLAG_INSTANCE_SPLIT, rt.RT2_FLAG_INSTANCES_INLINE


@pytest.mark.parametrize("name,extra", [(BOOK2, 0), ("cornell_box_scene_graph", rt.RT2_FLAG_NO_FLAT_EXTEND),
                                        ("cornell_box4", rt.RT2_FLAG_NO_FLAT_EXTEND), (BOOK2, rt.RT2_FLAG_GPU_LBVH)])
def test_all_instance_walks_report_the_same_hits(native_lib, name, extra):
    """The closest hit is an arg-min over the same leaves with the same model-space arithmetic in all three walks: the unified
    world tree, the two-pass split and the inline two-level walk must agree bit for bit (exact cross-space ties aside)."""
    scene = rt.Scene.load(scene_path(name))
    n_inst = len(scene.instances())
    assert 1 <= n_inst <= 4
    o, d, tm = fixed_rays(scene, 200000, seed=17)
    res = {}
    for label, flag, mode in (("unified", UNIFIED, 3), ("split", SPLIT, 2), ("inline", INLINE, 1)):
        tr = rt.RayTracer(scene, flags=extra | flag, dims=(32, 32))
        assert tr.stats()["instance_mode"] == mode, (label, tr.stats()["instance_mode"])
        res[label] = tr.intersect(o, d, tm, skip_media=True)
    a = res["unified"]
    hit = a["material"] >= 0
    assert hit.sum() > 50000 and (a["instance"][hit] >= 0).sum() > 500, "the rays must exercise the instances"
    for other in ("split", "inline"):
        b = res[other]
        assert np.array_equal(hit, b["material"] >= 0), other
        same = (_bits(a["t"]) == _bits(b["t"])) & (a["prim"] == b["prim"]) & (a["instance"] == b["instance"])
        assert (~same[hit]).sum() <= 2e-5 * hit.sum(), (other, int((~same[hit]).sum()))
        ok = hit & same
        for key in ("point", "normal"):
            assert np.array_equal(_bits(a[key][ok]), _bits(b[key][ok])), (other, key)
        assert np.array_equal(a["front_face"][ok], b["front_face"][ok])


def test_all_instance_walks_render_the_same_image(native_lib):
    dims, spp = (200, 200), 16
    scene = rt.Scene.load(scene_path(BOOK2), perlin_seed=5)
    kw = dict(num_samples=spp, max_depth=50, seed=77, dims=dims)
    img, rays, launches = {}, {}, {}
    for label, flag in (("unified", UNIFIED), ("split", SPLIT), ("inline", INLINE)):
        tr = rt.RayTracer(scene, flags=flag, **kw)
        tr.Update(spp)
        img[label] = tr.read_accum()
        rays[label], launches[label] = tr.stats()["rays"], tr.stats()["launches"]
    for other in ("split", "inline"):
        assert abs(rays["unified"] - rays[other]) <= 2e-5 * rays["unified"], (other, rays)
        same = np.all(_bits(img["unified"]) == _bits(img[other]), axis=-1)
        assert same.mean() > 0.999, (other, same.mean())  # identical hits -> identical paths (same Philox keys)
    assert launches["split"] > launches["inline"] == launches["unified"]  # the split adds one kernel per bounce


# ---- deferred Update -------------------------------------------------------------------------------------------------
def test_update_once_per_sample_is_batched_and_bit_identical(native_lib):
    """App.cpp:243-248 calls Update(scene) once per sample.  37 such calls = Update(37): same bits, and — because the frames
    are collected into wavefront batches — about the same number of kernel launches, not 37 times as many."""
    scene = rt.Scene.load(scene_path("cornell_original_test"))
    kw = dict(num_samples=64, max_depth=50, seed=11, dims=(96, 96), frames_per_batch=16)
    one = rt.RayTracer(scene, **kw)
    l0 = one.stats()["launches"]
    one.Update(37)
    a = one.read_accum()
    la = one.stats()["launches"] - l0
    many = rt.RayTracer(scene, **kw)
    l1 = many.stats()["launches"]
    for k in range(37):
        many.Update(1)
        assert many.FrameIdx() == k + 1  # frame_idx_++ per Update (RayTracer.cpp:61), traced or pending
    b = many.read_accum()
    lb = many.stats()["launches"] - l1
    assert np.array_equal(_bits(a), _bits(b))
    assert lb == la, (la, lb)
    st = many.stats()
    assert st["frames"] == 37 and st["pending_frames"] == 0 and st["paths"] == 96 * 96 * 37
    # pending frames are part of every read-out, and Reset drops them
    many.Update(3)
    assert many.FrameIdx() == 40
    m40 = many.NonConvertedPixels()
    one.Update(3)
    assert np.array_equal(_bits(m40), _bits(one.NonConvertedPixels()))
    many.Update(5)
    many.Reset()
    assert many.FrameIdx() == 0 and many.stats()["paths"] == 0
    many.Update(2)
    many.flush()
    assert many.stats()["frames"] == 2


# ---- several GPUs behind one handle ----------------------------------------------------------------------------------
def test_n_gpus_argument_is_validated(native_lib):
    scene = rt.Scene.load(scene_path("cornell_original_test"))
    n = native_lib.rt2_device_count()
    with pytest.raises(rt.Rt2Error):
        rt.RayTracer(scene, dims=(32, 32), n_gpus=n + 1)
    tr = rt.RayTracer(scene, dims=(32, 32), n_gpus=-1)  # every visible GPU
    assert tr.stats()["n_gpus"] == n
    tr.Update(3)
    assert tr.FrameIdx() == 3 and tr.stats()["paths"] == 32 * 32 * 3


@pytest.mark.parametrize("name,dims,spp", [(BOOK2, (160, 160), 37), ("cornell_original_test", (128, 128), 64)])
def test_multi_gpu_handle_equals_one_gpu(native_lib, name, dims, spp):
    """rt2_config.n_gpus = N: frames dealt round-robin to one replica per GPU inside the handle, read-out summed over peer
    memory in one kernel.  Same frames, same Philox keys => the image equals the 1-GPU image up to fp32 summation order."""
    n = native_lib.rt2_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    scene = rt.Scene.load(scene_path(name), perlin_seed=5)
    kw = dict(num_samples=spp, max_depth=50, seed=4242, dims=dims, frames_per_batch=8, flags=rt.RT2_FLAG_MOMENTS)
    one = rt.RayTracer(scene, **kw)
    one.Update(spp)
    ref_mean, ref_rgba = one.NonConvertedPixels(), one.Pixels()
    ref_s, ref_ss = one.read_accum(moments=True)
    for n_gpus in sorted({2, n}):
        multi = rt.RayTracer(scene, n_gpus=n_gpus, **kw)
        # an uneven call pattern: the partition must depend on the global frame index only
        done = 0
        for chunk in (1, 2, 5, spp - 8):
            multi.Update(chunk)
            done += chunk
            assert multi.FrameIdx() == done
        mean = multi.NonConvertedPixels()
        st = multi.stats()
        assert st["n_gpus"] == n_gpus and st["frames"] == spp and st["paths"] == dims[0] * dims[1] * spp
        assert st["rays"] == one.stats()["rays"]
        rel = np.abs(mean - ref_mean).max() / max(float(np.abs(ref_mean).max()), 1e-9)
        assert rel <= 1e-6, rel
        rgba = multi.Pixels()
        assert (rgba != ref_rgba).mean() < 1e-4  # a last-bit difference of the mean may cross an 8-bit boundary
        s, ss = multi.read_accum(moments=True)
        assert np.allclose(s, ref_s, rtol=2e-6, atol=1e-6) and np.allclose(ss, ref_ss, rtol=2e-6, atol=1e-6)
        # a second read-out and more frames afterwards (nothing is reduced in place)
        assert np.array_equal(_bits(multi.NonConvertedPixels()), _bits(mean))
        multi.Update(3)
        one2 = rt.RayTracer(scene, **kw)
        one2.Update(spp + 3)
        m2 = multi.NonConvertedPixels()
        assert np.abs(m2 - one2.NonConvertedPixels()).max() / max(float(np.abs(ref_mean).max()), 1e-9) <= 1e-6
        # checkpoint across the handle: sums out, sums back in, continue
        s, ss = multi.read_accum(moments=True)
        again = rt.RayTracer(scene, n_gpus=n_gpus, **kw)
        again.write_accum(s, ss, spp + 3)
        assert again.FrameIdx() == spp + 3
        again.Update(5)
        multi.Update(5)
        assert np.abs(again.NonConvertedPixels() - multi.NonConvertedPixels()).max() <= 1e-6 * max(float(np.abs(ref_mean).max()), 1e-9)
        # single-GPU plumbing is refused on a multi-GPU handle
        with pytest.raises(rt.Rt2Error):
            multi.accum_device_ptr()


def test_raytrace_2_binary_uses_every_gpu(native_lib, tmp_path):
    """`raytrace_2 <scene> out.png` (App.cpp:81-249): settings.json, argv, PNG — on all GPUs of the box by default, --gpus / --spp
    overrides; two runs with the same --seed on different GPU counts give (nearly) the same PNG."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "raytrace2_b200", "bin", "raytrace_2")
    root = tmp_path / "root"
    (root / "local" / "data").mkdir(parents=True)
    os.symlink(os.path.join(ROOT, "data"), root / "data")
    (root / "local" / "data" / "settings.json").write_text(json.dumps(
        {"num_samples": 400, "render_once": True, "save_after_render_once": True, "max_depth": 50, "render_window": False}))
    env = dict(os.environ, RAYTRACE2_ROOT=str(root))
    outs = []
    n = native_lib.rt2_device_count()
    for gpus in ([1, n] if n > 1 else [1]):
        out = tmp_path / f"o{gpus}.png"
        cmd = [exe, "data/cornell_original_test", str(out), "--spp", "36", "--seed", "9"] + (["--gpus", str(gpus)] if gpus != n else [])
        p = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "Num Samples: 36" in p.stdout and f"on {gpus} GPU(s)" in p.stdout
        from PIL import Image
        outs.append(np.asarray(Image.open(out).convert("RGB"), np.int32))
        assert outs[-1].shape == (600, 600, 3)
    if len(outs) == 2:
        assert (np.abs(outs[0] - outs[1]) > 1).mean() < 1e-4
    p = subprocess.run([exe, "data/cornell_original_test", str(tmp_path / "x.png"), "--gpus", str(n + 1)], cwd=root, env=env,
                       capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "--gpus" in p.stderr


# ---- device textures at fixed points ---------------------------------------------------------------------------------
def test_checker_texture_matches_reference_bit_for_bit(native_lib, tmp_path):
    g = np.load(os.path.join(GOLDEN, "texture_checker.npz"))
    path = tmp_path / "scene.json"
    path.write_bytes(g["scene_json"].tobytes())
    scene = rt.Scene.load(str(path), data_dir=os.path.dirname(scene_path("x")))
    tr = rt.RayTracer(scene, dims=(16, 16))
    got = tr.texture_value(0, g["points"])
    assert np.array_equal(got, g["value_0"])  # checker parity: ivec3(floor(p / scale)), sum % 2 (Texture.cpp:7-14)
    assert len(np.unique(got, axis=0)) == 2


def test_noise_textures_match_reference_at_fixed_points(native_lib, tmp_path):
    """Marble and Perlin (Texture.cpp:16-22, PerlinNoiseGen.cpp:52-88) with the reference's own tables: <= 1e-5 relative to the
    texture's range.  The only differences are FMA contraction in the interpolation and sinf vs libm's sin."""
    g = np.load(os.path.join(GOLDEN, "texture_noise.npz"))
    path = tmp_path / "scene.json"
    path.write_bytes(g["scene_json"].tobytes())
    scene = rt.Scene.load(str(path), data_dir=os.path.dirname(scene_path("x")))
    tex = sorted(int(k.split("_")[1]) for k in g.files if k.startswith("value_"))
    assert scene.desc.n_perlin == len(tex)
    types = {int(t["noise_type"]) for t in scene.textures() if int(t["type"]) == 2}
    assert types == {0, 1}, "both Perlin (0) and marble (1) must be covered"
    for slot, ti in enumerate(tex):
        pi = int(scene.textures()[ti]["perlin_idx"])
        assert pi == slot
        scene.set_perlin(pi, g[f"perm_x_{ti}"], g[f"perm_y_{ti}"], g[f"perm_z_{ti}"], g[f"vec_{ti}"])
    tr = rt.RayTracer(scene, dims=(16, 16))
    for ti in tex:
        got = tr.texture_value(ti, g["points"])
        ref = g[f"value_{ti}"]
        scale = float(np.abs(ref).max())
        assert scale > 0.3
        err = np.abs(got - ref).max() / scale
        assert err <= 1e-5, (ti, err)


# ---- exact ties -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, rt.RT2_FLAG_NO_FLAT_EXTEND, rt.RT2_FLAG_NO_FLAT_EXTEND | rt.RT2_FLAG_GPU_LBVH])
def test_coincident_quads_resolve_like_the_reference(native_lib, port_oracle, tmp_path, flags):
    """Two boxes sharing a face plane + a quad lying in the floor plane: rays that hit coincident faces must report the
    primitive the reference reports — the LAST one in its BVH / list order (closed interval, HittableList.cpp:8-22)."""
    doc = {"camera": {"fov": 40, "center": [0, 2, 9], "look_at": [0, 1, 0], "width": 64, "aspect_ratio": 1.0},
           "materials": [{"type": "lambertian", "albedo": [0.1 * (i + 1)] * 3} for i in range(6)],
           "primitives": [{"type": "box", "a": [-2, 0, -1], "b": [0, 2, 1], "material": 0},
                          {"type": "box", "a": [0, 0, -1], "b": [2, 1, 1], "material": 1},
                          {"type": "quad", "q": [-5, 0, -5], "u": [10, 0, 0], "v": [0, 0, 10], "material": 2},
                          {"type": "quad", "q": [-1, 0, -1], "u": [2, 0, 0], "v": [0, 0, 2], "material": 3},
                          {"type": "box", "a": [-2, 2, -1], "b": [0, 3, 1], "material": 4},
                          {"type": "quad", "q": [-1, 0, -1], "u": [2, 0, 0], "v": [0, 0, 2], "material": 5}],
           "scene": [{"primitive": i} for i in range(6)]}
    path = tmp_path / "ties.json"
    path.write_text(json.dumps(doc))
    scene = rt.Scene.load(str(path))
    tr = rt.RayTracer(scene, flags=flags, dims=(16, 16))
    port = port_oracle.PortScene(str(path), 16)
    rng = np.random.default_rng(3)
    n = 40000
    # rays aimed at the shared planes x = 0 (between the boxes), y = 0 (floor, box bottoms, the two coincident quads), y = 2
    target = np.stack([rng.uniform(-2, 2, n), rng.uniform(0, 2.5, n), rng.uniform(-1, 1, n)], axis=1)
    which = rng.integers(0, 3, n)
    target[which == 0, 0] = 0.0
    target[which == 1, 1] = 0.0
    target[which == 2, 1] = 2.0
    origin = target + rng.normal(size=(n, 3)) * 1.5
    d = (target - origin)
    d *= rng.uniform(0.3, 1.7, (n, 1))
    o, d = origin.astype(np.float32), d.astype(np.float32)
    g = tr.intersect(o, d)
    r = port.intersect(o, d)
    hit = r["hit"].astype(bool)
    assert np.array_equal(g["material"] >= 0, hit)
    assert np.array_equal(_bits(g["t"][hit]), _bits(r["t"][hit]))
    # every primitive has its own material, so the material identifies the winner of a tie
    assert np.array_equal(g["material"][hit], r["material"][hit])
    assert np.array_equal(g["normal"][hit], r["normal"][hit]) and np.array_equal(g["front_face"][hit], r["front_face"][hit].astype(np.uint32))
    # and there really are ties in this set: of the two coincident quads (materials 3 and 5) only the one the reference visits
    # later can ever be reported, and the shared plane x = 0 is hit from inside the boxes
    assert min((g["material"] == 3).sum(), (g["material"] == 5).sum()) == 0
    on_floor = hit & (g["point"][:, 1] == 0) & (np.abs(g["point"][:, 0]) < 1) & (np.abs(g["point"][:, 2]) < 1)
    assert on_floor.sum() > 100 and (hit & (g["point"][:, 0] == 0)).sum() > 100


# ---- legacy scenes against the real reference -------------------------------------------------------------------------
LEGACY_GOLDEN = ["final_render_book_1", "light_scene1", "checker_test", "cornell_box2"]


@pytest.mark.parametrize("name", LEGACY_GOLDEN)
def test_legacy_scene_hits_match_reference_golden(native_lib, name):
    """hits_<legacy>.npz come from the UNMODIFIED reference run on the converted file (tests/golden/make_golden_r2.py); the
    product loads the ORIGINAL legacy file through its adapter."""
    g = np.load(os.path.join(GOLDEN, f"hits_{name}.npz"))
    tr = rt.RayTracer(rt.Scene.load(scene_path(name)), dims=(16, 16))
    got = tr.intersect(g["origins"], g["directions"], g["times"], skip_media=True)
    hit = g["hit"].astype(bool)
    assert hit.sum() > 500
    assert np.array_equal(got["material"] >= 0, hit)
    assert np.array_equal(_bits(got["t"][hit]), _bits(g["t"][hit]))
    assert np.array_equal(_bits(got["point"][hit]), _bits(g["point"][hit]))
    assert np.array_equal(got["normal"][hit], g["normal"][hit])
    assert np.array_equal(got["material"][hit], g["material"][hit])
    assert np.array_equal(got["front_face"][hit], g["front_face"][hit].astype(np.uint32))


@pytest.mark.parametrize("name", ["final_render_book_1", "light_scene1", "cornell_box2"])
def test_legacy_scene_render_matches_reference_golden_moments(native_lib, name):
    g = np.load(os.path.join(GOLDEN, f"moments_{name}.npz"))
    dims, spp = tuple(int(x) for x in g["dims"]), int(g["spp"])
    scene = rt.Scene.load(scene_path(name))
    if "perm_x" in g.files:
        pi = int(scene.textures()[int(g["noise_tex"])]["perlin_idx"])
        scene.set_perlin(pi, g["perm_x"], g["perm_y"], g["perm_z"], g["vec"])
    tr = rt.RayTracer(scene, num_samples=spp, max_depth=50, seed=99, flags=rt.RT2_FLAG_MOMENTS, dims=dims)
    tr.Update(spp)
    s, ss = tr.read_accum(moments=True)
    z, valid = parity.z_scores(s, ss, spp, g["sum"].astype(np.float64), g["sumsq"].astype(np.float64), spp)
    st = parity.summary(z, valid)
    assert st["n"] > 1000 and abs(st["mean_z"]) < 4.0 / np.sqrt(st["n"]) + 0.01, st
    assert 0.90 < st["std_z"] < 1.06 and st["frac_gt3"] < 0.005, st
    rpp = tr.stats()["rays"] / (dims[0] * dims[1] * spp)
    assert abs(rpp - float(g["rays"]) / (dims[0] * dims[1] * spp)) < 0.01 * rpp + 0.01


# ---- BASELINE config 1 at full size -------------------------------------------------------------------------------------
def test_cornell_full_size_1024spp_against_reference_tiles(native_lib):
    """600 x 600, 1024 spp, max_depth 50 (SURVEY §8d C1 parity configuration) against the reference's own render, compared on
    4 x 4-pixel tiles: z = (tile sums differ) / sqrt(sum of the per-pixel variances of both renders)."""
    g = np.load(os.path.join(GOLDEN, "tiles_cornell_original_test_600_1024.npz"))
    tile, spp, dims = int(g["tile"]), int(g["spp"]), tuple(int(x) for x in g["dims"])
    scene = rt.Scene.load(scene_path("cornell_original_test"))
    tr = rt.RayTracer(scene, num_samples=spp, max_depth=50, seed=31337, flags=rt.RT2_FLAG_MOMENTS, dims=dims)
    for _ in range(spp):
        tr.Update(1)  # the reference's call pattern
    s, ss = (a.astype(np.float64) for a in tr.read_accum(moments=True))

    def tiles(a):
        h, w, c = a.shape
        return a.reshape(h // tile, tile, w // tile, tile, c).sum(axis=(1, 3))
    n = float(spp)
    # per pixel: var of the mean = (sumsq - sum^2 / n) / (n (n - 1)); a tile's mean-sum has the sum of those
    var_a = (tiles(ss) - tiles(s * s) / n) / (n * (n - 1))
    var_b = (g["sumsq"].astype(np.float64) - g["sqsum"].astype(np.float64) / n) / (n * (n - 1))
    diff = (tiles(s) - g["sum"].astype(np.float64)) / n
    var = var_a + var_b
    valid = var > 1e-12
    z = diff[valid] / np.sqrt(var[valid])
    assert z.size > 55000  # 150 x 150 tiles x 3 channels, minus the black ones (zero variance in both renders)
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size) + 0.01, z.mean()
    assert 0.90 < z.std() < 1.06, z.std()
    assert (np.abs(z) > 3).mean() < 0.005
    rpp = tr.stats()["rays"] / (dims[0] * dims[1] * spp)
    assert abs(rpp - float(g["rays"]) / (dims[0] * dims[1] * spp)) < 0.005 * rpp, rpp


# ---- the traversal stack cannot overflow: tree depths are verified at upload ------------------------------------------------
def test_tree_depths_fit_the_traversal_stack(native_lib):
    """The walks keep one stack entry per tree level and do no bounds checks; rt2_create / rt2_upload_scene compute the depth of
    every tree on the device (host SAH trees and device LBVH trees alike) and refuse a scene that does not fit."""
    for name, flags in [(BOOK2, 0), (BOOK2, INLINE), (BOOK2, SPLIT), (BOOK2, rt.RT2_FLAG_GPU_LBVH), (BOOK2, rt.RT2_FLAG_GPU_LBVH | INLINE),
                        ("final_render_book_1", rt.RT2_FLAG_GPU_LBVH), ("cornell_box4", rt.RT2_FLAG_NO_FLAT_EXTEND)]:
        tr = rt.RayTracer(rt.Scene.load(scene_path(name)), num_samples=4, dims=(160, 90), flags=flags)
        tr.Update(4)
        tr.NonConvertedPixels()
        st = tr.stats()
        assert 1 <= st["max_stack_need"] <= 63 and st["stack_overflows"] == 0, (name, flags, st["max_stack_need"])
    # the inline walk nests TLAS + sentinel + BLAS on one stack; the split walks need only the deeper of the two trees
    scene = rt.Scene.load(scene_path(BOOK2))
    need = {f: rt.RayTracer(scene, dims=(32, 32), flags=f).stats()["max_stack_need"] for f in (UNIFIED, SPLIT, INLINE)}
    assert need[INLINE] > need[SPLIT] and need[INLINE] > need[UNIFIED]


def test_duplicate_keys_do_not_make_a_deep_lbvh(native_lib):
    """50 000 spheres, device LBVH with 63-bit Morton keys: duplicates are split by index bits, so the tree stays shallow."""
    scene = rt.Scene.synthetic_spheres(50000, seed=5, width=64, height=36, host_bvh=False)
    tr = rt.RayTracer(scene, num_samples=4, flags=rt.RT2_FLAG_GPU_LBVH)
    tr.Update(4)
    tr.NonConvertedPixels()
    st = tr.stats()
    assert st["stack_overflows"] == 0 and st["rays"] > 0 and 10 <= st["max_stack_need"] <= 40, st["max_stack_need"]


# ---- 32-byte quantised node pairs (rt_qnodes.cu) against the 64-byte float pairs -------------------------------------------------
@pytest.mark.parametrize("name,extra", [(BOOK2, 0), (BOOK2, rt.RT2_FLAG_INSTANCES_INLINE), ("cornell_box_scene_graph", rt.RT2_FLAG_NO_FLAT_EXTEND),
                                        ("cornell_box_scene_graph", rt.RT2_FLAG_NO_FLAT_EXTEND | rt.RT2_FLAG_INSTANCES_INLINE),
                                        (BOOK2, rt.RT2_FLAG_GPU_LBVH), ("cornell_vol_test", rt.RT2_FLAG_NO_FLAT_EXTEND)])
def test_compact_nodes_report_the_same_hits(native_lib, name, extra):
    """Box tests only cull: with every quantised box containing its float box (outward rounding + one cell), the closest hit of
    every ray — t, primitive, instance, point, normal — must be bit-identical with the float-node walk."""
    scene = rt.Scene.load(scene_path(name))
    o, d, tm = fixed_rays(scene, 200000, seed=23)
    tq = rt.RayTracer(scene, flags=extra, dims=(32, 32))
    tf = rt.RayTracer(scene, flags=extra | rt.RT2_FLAG_FLOAT_NODES, dims=(32, 32))
    sq, sf = tq.stats(), tf.stats()
    assert sf["compact_nodes"] == 0
    assert sq["compact_nodes"] == 1, (name, sq["node_inflation"])
    assert 0.0 <= sq["node_inflation"] <= 0.05
    a, b = tq.intersect(o, d, tm, skip_media=True), tf.intersect(o, d, tm, skip_media=True)
    hit = a["material"] >= 0
    assert hit.sum() > 20000
    for key in ("t", "point", "normal"):  # (compact-node test)
        assert np.array_equal(_bits(a[key]), _bits(b[key])), key
    for key in ("prim", "instance", "material", "front_face"):
        assert np.array_equal(a[key], b[key]), key


def test_compact_nodes_render_the_same_image(native_lib):
    dims, spp = (200, 200), 16
    scene = rt.Scene.load(scene_path(BOOK2), perlin_seed=5)
    kw = dict(num_samples=spp, max_depth=50, seed=77, dims=dims)
    img, work = {}, {}
    for label, flag in (("compact", 0), ("float", rt.RT2_FLAG_FLOAT_NODES)):
        tr = rt.RayTracer(scene, flags=flag, **kw)
        tr.set_profiling(True)  # counts node visits
        tr.Update(spp)
        img[label] = tr.read_accum()
        st = tr.stats()
        assert st["compact_nodes"] == (1 if label == "compact" else 0)
        work[label] = st["box_pair_tests"]
    assert np.array_equal(_bits(img["compact"]), _bits(img["float"]))
    # the padded boxes may be entered a little more often, never dramatically
    assert work["float"] > 0 and work["float"] <= work["compact"] <= 1.05 * work["float"], work


@pytest.mark.parametrize("name", ["final_render_book_1", "synthetic"])
def test_scenes_with_small_leaves_keep_float_nodes(native_lib, name):
    """Book 1's final scene and BASELINE config 5: spheres of radius 0.2 in a world spanned by a ground sphere of radius 1000 —
    one 15-bit grid cell (0.06) is a sizeable fraction of a leaf.  The mean growth of the boxes' surface area exceeds the limit
    and the renderer keeps the float nodes (measured: the quantised walk visits 11 % / 290 % more nodes there)."""
    if name == "synthetic":
        scene, flags = rt.Scene.synthetic_spheres(200000, seed=3, host_bvh=False), rt.RT2_FLAG_GPU_LBVH
    else:
        scene, flags = rt.Scene.load(scene_path(name)), 0
    st = rt.RayTracer(scene, flags=flags, dims=(64, 64)).stats()
    assert st["node_inflation"] > 0.05 and st["compact_nodes"] == 0, st["node_inflation"]
