#!/usr/bin/env python3
"""Round-2 pins from the REAL reference (oracle/_ref/libref_oracle.so), added without touching the round-1 files:

  hits_<legacy scene>.npz, moments_<legacy scene>.npz
        Legacy-format scene files make HEAD's loader throw, so they are first rewritten into the current format by
        tools/convert_legacy.py (into a temp dir) — the unmodified reference then loads and renders them.  Pins BASELINE
        config 2 (final_render_book_1) and legacy quad / light / checker / box scenes to the reference itself.
  texture_<scene>.npz
        ref Texture::Value at fixed points for checker / marble / perlin textures (+ the Perlin tables used), the pin of the
        device texture code (rt2_texture_value).
  tiles_cornell_original_test_600_1024.npz
        BASELINE config 1 at FULL size (600 x 600, 1024 spp): per 4x4-pixel tile the sums over its pixels of
        (sum, sum of squares, squared sum) of the reference render — enough for a tile z-test, 30x smaller than per-pixel moments.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_r2.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from convert_legacy import convert_file  # noqa: E402
from make_golden import N_RAYS  # noqa: E402
from oracle.ref_oracle import RefScene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LEGACY = ["final_render_book_1", "light_scene1", "checker_test", "cornell_box2"]
LEGACY_MOMENTS = [("final_render_book_1", (96, 54), 144), ("light_scene1", (80, 45), 256), ("cornell_box2", (60, 60), 256)]
TILE = 4


def rays_for_legacy(ref, name, seed):
    """Half camera-like rays from around the camera centre towards the scene, half interior rays, half of each non-unit."""
    rng = np.random.default_rng(seed)
    cam = ref.camera()
    center = cam[0:3].astype(np.float64)
    if name.startswith("cornell"):
        lo, hi = np.array([-60.0, -60.0, -850.0]), np.array([620.0, 620.0, 620.0])
    else:
        lo, hi = np.array([-14.0, -1.0, -14.0]), np.array([14.0, 6.0, 14.0])
    n = N_RAYS
    o = rng.uniform(lo, hi, size=(n, 3))
    o[: n // 3] = center + rng.normal(scale=0.3, size=(n // 3, 3))
    v = rng.normal(size=(n, 3))
    v[: n // 3] = rng.uniform(lo, hi, size=(n // 3, 3)) * 0.5 - o[: n // 3]
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    scale = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(0.05, 2.0, n))
    return o.astype(np.float32), (v * scale[:, None]).astype(np.float32), rng.random(n).astype(np.float32)


def tile_sums(a, tile):
    h, w, c = a.shape
    return a.reshape(h // tile, tile, w // tile, tile, c).sum(axis=(1, 3))


def main():
    td = tempfile.mkdtemp()
    conv = {}
    for name in LEGACY:
        conv[name] = os.path.join(td, name + ".json")
        convert_file(os.path.join(ROOT, "data", name + ".json"), conv[name], os.path.join(ROOT, "data"))
    for name in LEGACY:
        ref = RefScene(conv[name], 16)
        o, d, t = rays_for_legacy(ref, name, 123)
        r0 = ref.intersect(o, d, t)
        np.savez_compressed(os.path.join(HERE, f"hits_{name}.npz"), origins=o, directions=d, times=t, hit=r0["hit"], t=r0["t"],
                            point=r0["point"], normal=r0["normal"], front_face=r0["front_face"], material=r0["material"],
                            deterministic=np.ones(o.shape[0], bool))
        print(name, "rays", o.shape[0], "hits", int(r0["hit"].sum()))
    for name, dims, spp in LEGACY_MOMENTS:
        ref = RefScene(conv[name], spp, dims=dims)
        extra = {}
        doc = json.load(open(conv[name]))
        noise = [i for i, tx in enumerate(doc.get("textures") or []) if tx.get("type") == "noise"]
        if noise:
            px = ref.perlin_get(noise[0])
            extra = dict(perm_x=px[0], perm_y=px[1], perm_z=px[2], vec=px[3], noise_tex=noise[0])
        s, ss, rays, sec = ref.render(0, spp, 50, 0, True)
        np.savez_compressed(os.path.join(HERE, f"moments_{name}.npz"), sum=s.astype(np.float32), sumsq=ss.astype(np.float32),
                            spp=spp, dims=np.array(dims), rays=rays, **extra)
        print(name, dims, spp, "rays/path", rays / (dims[0] * dims[1] * spp), f"{sec:.1f}s")
    # texture values at fixed points: checker (checker_test), marble (light_scene1), perlin + marble (book 2 has marble only;
    # a perlin-type copy of its texture is added to the converted light scene)
    rng = np.random.default_rng(77)
    pts = rng.uniform(-12, 12, size=(4096, 3)).astype(np.float32)
    pts[:64] = np.round(pts[:64])  # lattice points: the checker's floor() boundaries and Perlin's integer cells
    doc = json.load(open(conv["light_scene1"]))
    doc["textures"].append({"type": "noise", "scale": 2.5, "noise_type": 0, "albedo": [0.9, 0.5, 0.2]})
    both = os.path.join(td, "light_scene1_perlin.json")
    json.dump(doc, open(both, "w"))
    for tag, path, tex in [("checker", conv["checker_test"], [0]), ("noise", both, [0, len(doc["textures"]) - 1])]:
        ref = RefScene(path, 16)
        out = {"points": pts}
        for ti in tex:
            out[f"value_{ti}"] = ref.texture_value(ti, pts)
            if tag == "noise":
                px = ref.perlin_get(ti)
                out.update({f"perm_x_{ti}": px[0], f"perm_y_{ti}": px[1], f"perm_z_{ti}": px[2], f"vec_{ti}": px[3]})
        out["scene_json"] = np.frombuffer(open(path, "rb").read(), np.uint8)
        np.savez_compressed(os.path.join(HERE, f"texture_{tag}.npz"), **out)
        print("texture", tag, tex)
    # BASELINE config 1 at full size
    name, dims, spp = "cornell_original_test", (600, 600), 1024
    ref = RefScene(os.path.join(ROOT, "data", name + ".json"), spp, dims=dims)
    s, ss, rays, sec = ref.render(0, spp, 50, 0, True)
    s64, ss64 = s.astype(np.float64), ss.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, f"tiles_{name}_600_1024.npz"), tile=TILE, spp=spp, dims=np.array(dims), rays=rays,
                        sum=tile_sums(s64, TILE).astype(np.float32), sumsq=tile_sums(ss64, TILE).astype(np.float32),
                        sqsum=tile_sums(s64 * s64, TILE).astype(np.float32))
    print(name, dims, spp, "rays/path", rays / (dims[0] * dims[1] * spp), f"{sec:.1f}s")


if __name__ == "__main__":
    main()
