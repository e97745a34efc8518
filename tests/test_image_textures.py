"""Image textures (schema extension; include/rt2.h RT2_TEX_IMAGE): the host decoder (PNG / PPM) against PIL,
the lookup rule in the CPU restatement, and — on the GPU — HitRecord::uv and rendered images against the restatement."""
import json
import os

import numpy as np
import pytest

import raytrace2_b200 as rt
from _textures import test_picture as make_picture, textured_scene
from raytrace2_b200 import scene_builder as sb


def _lin(a):
    return (a.astype(np.float32) / np.float32(255.0)) ** np.float32(2.2)


def _scene_with(path, data_dir):
    b = sb.SceneBuilder(width=32)
    b.place(b.sphere((0, 0, 0), 1, b.textured(b.image(path))))
    return rt.Scene.from_builder(b, data_dir=str(data_dir))


@pytest.mark.parametrize("mode,ext", [("RGB", "png"), ("RGBA", "png"), ("L", "png"), ("P", "png"), ("RGB", "ppm")])
def test_decoder_matches_pil(native_lib, tmp_path, mode, ext):
    from PIL import Image
    pic = make_picture()
    im = Image.fromarray(pic, "RGB").convert(mode)
    name = f"pic_{mode}.{ext}"
    im.save(tmp_path / name)
    expect = np.asarray(Image.open(tmp_path / name).convert("RGB"), np.uint8)
    got = _scene_with(name, tmp_path).images()
    assert len(got) == 1 and got[0].shape == (pic.shape[0], pic.shape[1], 4)
    assert np.allclose(got[0][..., :3], _lin(expect), rtol=2e-6, atol=1e-7)
    assert np.all(got[0][..., 3] == 1.0)


def test_ascii_ppm_and_missing_file(native_lib, tmp_path):
    pic = make_picture(5, 3)
    with open(tmp_path / "a.ppm", "w") as f:
        f.write("P3\n# comment\n5 3\n255\n" + " ".join(str(int(v)) for v in pic.reshape(-1)) + "\n")
    got = _scene_with("a.ppm", tmp_path).images()[0]
    assert np.allclose(got[..., :3], _lin(pic), rtol=2e-6, atol=1e-7)
    cyan = _scene_with("does_not_exist.png", tmp_path).images()[0]
    assert cyan.shape == (1, 1, 4) and np.array_equal(cyan[0, 0, :3], [0, 1, 1])


def test_truncated_and_out_of_range_binary_ppm(native_lib, tmp_path):
    """A P6 file that ends right after its header (or short of pixels) must be rejected, not read out of bounds; 16-bit samples
    above maxval are clamped like in the ASCII path."""
    (tmp_path / "hdr_only.ppm").write_bytes(b"P6 4 4 255")
    (tmp_path / "short.ppm").write_bytes(b"P6\n4 4\n255\n" + bytes(10))
    for name in ("hdr_only.ppm", "short.ppm"):
        got = _scene_with(name, tmp_path).images()[0]  # undecodable -> the 1x1 cyan debugging texture, like a missing file
        assert got.shape == (1, 1, 4) and np.array_equal(got[0, 0, :3], [0, 1, 1])
    # maxval 1000, one sample of 2000 (> maxval) and one of 500
    body = b"".join(int(v).to_bytes(2, "big") for v in (2000, 500, 0))
    (tmp_path / "clamp.ppm").write_bytes(b"P6\n1 1\n1000\n" + body)
    got = _scene_with("clamp.ppm", tmp_path).images()[0]
    assert got.shape == (1, 1, 4)
    assert np.allclose(got[0, 0, :3], _lin(np.array([255, 128, 0], np.uint8)), rtol=2e-6, atol=1e-7)


def test_restatement_lookup_rule(port_oracle, tmp_path):
    from PIL import Image
    pic = make_picture()
    Image.fromarray(pic, "RGB").save(tmp_path / "pic.png")
    b = textured_scene(sb)
    (tmp_path / "scene.json").write_text(b.to_json())
    port = port_oracle.PortScene(str(tmp_path / "scene.json"), 16)
    rng = np.random.default_rng(1)
    uv = rng.uniform(-0.2, 1.2, size=(5000, 2)).astype(np.float32)
    got = port.texture_value_uv(0, np.zeros((5000, 3), np.float32), uv)
    h, w = pic.shape[:2]
    u = np.clip(uv[:, 0], 0, 1)
    v = np.float32(1.0) - np.clip(uv[:, 1], 0, 1)
    i = np.minimum((u * np.float32(w)).astype(np.int64), w - 1)
    j = np.minimum((v * np.float32(h)).astype(np.int64), h - 1)
    assert np.allclose(got, _lin(pic)[j, i], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_gpu_uv_matches_restatement(native_lib, port_oracle, tmp_path):
    from PIL import Image
    from _rays import fixed_rays
    Image.fromarray(make_picture(), "RGB").save(tmp_path / "pic.png")
    b = textured_scene(sb)
    (tmp_path / "scene.json").write_text(b.to_json())
    scene = rt.Scene.load(str(tmp_path / "scene.json"))
    port = port_oracle.PortScene(str(tmp_path / "scene.json"), 16)
    o, d, t = fixed_rays(scene, 40000, seed=4)
    g = rt.RayTracer(scene, dims=(32, 32)).intersect(o, d, t)
    hit, uv = port.intersect_uv(o, d, t)
    assert np.array_equal(g["material"] >= 0, hit.astype(bool)) and hit.sum() > 5000
    m = hit.astype(bool)
    gu = (g["uv16"][m] & 0xFFFF).astype(np.float32) / 65535.0
    gv = (g["uv16"][m] >> 16).astype(np.float32) / 65535.0
    eu, ev = np.clip(uv[m, 0], 0, 1), np.clip(uv[m, 1], 0, 1)
    du = np.minimum(np.abs(gu - eu), 1.0 - np.abs(gu - eu))  # the sphere's u wraps at the seam
    assert du.max() < 3e-5 and np.abs(gv - ev).max() < 3e-5


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, rt.RT2_FLAG_NO_FUSED_SHADE])
def test_gpu_textured_render_matches_restatement(native_lib, port_oracle, tmp_path, flags):
    from PIL import Image
    from raytrace2_b200 import parity
    Image.fromarray(make_picture(), "RGB").save(tmp_path / "pic.png")
    b = textured_scene(sb)
    (tmp_path / "scene.json").write_text(b.to_json())
    spp, dims = 144, (160, 160)
    scene = rt.Scene.load(str(tmp_path / "scene.json"))
    port = port_oracle.PortScene(str(tmp_path / "scene.json"), spp, dims=dims)
    tracer = rt.RayTracer(scene, num_samples=spp, max_depth=50, seed=11, flags=rt.RT2_FLAG_MOMENTS | flags, dims=dims)
    tracer.Update(spp)
    s, ss = tracer.read_accum(moments=True)
    rs, rss, _, _ = port.render(0, spp, 50, 0, True)
    z, valid = parity.z_scores(s, ss, spp, rs, rss, spp)
    st = parity.summary(z, valid)
    assert st["n"] > 1000 and abs(st["mean_z"]) < 4.0 / np.sqrt(st["n"]) + 0.01, st
    # most pixels see a texture lit by the constant background: their variance is the stratified sub-pixel jitter
    # (RayTracer.cpp:57-60), which s^2/n overestimates -> z is under-dispersed; a mismatch would show as std(z) > 1
    assert 0.5 < st["std_z"] < 1.06 and st["frac_gt3"] < 0.005, st
    tz = parity.tile_z_scores(s, ss, spp, rs, rss, spp, tile=20)
    assert np.abs(tz).max() < 4.5


def test_unsupported_png_variants_fall_back_to_cyan_with_a_warning(native_lib, tmp_path):
    """16-bit and interlaced PNGs are outside the decoder's subset: the texture becomes the 1x1 cyan placeholder (the scene
    still loads), like a missing file."""
    from PIL import Image
    pic = make_picture(16, 8)
    Image.fromarray(pic[..., 0].astype(np.uint16) * 257).save(tmp_path / "deep.png")
    Image.fromarray(pic, "RGB").save(tmp_path / "ok.png")
    # flip the interlace byte of a valid file (IHDR data byte 12) and fix nothing else: the decoder must reject it on the flag
    raw = bytearray((tmp_path / "ok.png").read_bytes())
    raw[8 + 8 + 12] = 1
    (tmp_path / "interlaced.png").write_bytes(bytes(raw))
    (tmp_path / "garbage.png").write_bytes(b"not an image at all")
    for name in ("deep.png", "interlaced.png", "garbage.png"):
        img = _scene_with(name, tmp_path).images()[0]
        assert img.shape == (1, 1, 4) and np.array_equal(img[0, 0, :3], [0, 1, 1]), name
    ok = _scene_with("ok.png", tmp_path).images()[0]
    assert ok.shape == (8, 16, 4)


def test_image_texture_through_a_checker(native_lib, port_oracle, tmp_path):
    """An image texture as one of a checker's children: the (u, v) of the hit travel through the recursion (Texture.cpp:7-11)."""
    from PIL import Image
    pic = make_picture(32, 16)
    Image.fromarray(pic, "RGB").save(tmp_path / "pic.png")
    b = sb.SceneBuilder(width=32)
    tex = b.checker(2.0, b.image("pic.png"), b.solid((0.1, 0.2, 0.3)))
    b.place(b.sphere((0, 0, 0), 1.0, b.textured(tex)))
    (tmp_path / "scene.json").write_text(b.to_json())
    port = port_oracle.PortScene(str(tmp_path / "scene.json"), 4)
    pts = np.float32([[0.5, 0.5, 0.5], [2.5, 0.5, 0.5], [-0.5, 0.5, 0.5]])   # floor(p / 2) parity: even, odd, odd
    uv = np.float32([[0.25, 0.75], [0.25, 0.75], [0.9, 0.1]])
    got = port.texture_value_uv(2, pts, uv)
    h, w = pic.shape[:2]
    i, j = min(int(0.25 * w), w - 1), min(int((1 - 0.75) * h), h - 1)
    assert np.allclose(got[0], _lin(pic)[j, i], rtol=1e-6)
    assert np.allclose(got[1], [0.1, 0.2, 0.3]) and np.allclose(got[2], [0.1, 0.2, 0.3])
