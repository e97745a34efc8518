#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libref_oracle.so = the unmodified sources compiled
by oracle/Makefile).  The reference has no tests or golden vectors of its own, so these are the pins:

  hits_<scene>.npz    fixed rays + the reference's closest-hit records (hit, t, point, normal, front_face, material);
                      for scenes with constant media (stochastic Hit) a `deterministic` mask marks rays whose record
                      cannot have been influenced by a medium (see below)
  camera_<scene>.npy  the reference's camera block after Camera::Update for the authored dims
  span1_<scene>.npy   which top-level objects sit in span-1 BVH leaves (quirk Q2)
  moments_<scene>.npz per-pixel sum / sum of squares of a low-resolution reference render (statistical pin)
  tonemap.npz         util::WriteImage's 8-bit conversion on a fixed float image (decoded from the PNG it wrote)

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import struct
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_oracle import RefScene, write_image  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = ["cornell_original_test", "cornell_box_scene_graph", "cornell_box4", "cornell_volume_10000_samples",
          "book2_final_scene_10000_samples"]
N_RAYS = 6000


def rays_for(scene_name, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-60.0, -60.0, -850.0]), np.array([620.0, 620.0, 620.0])
    if scene_name.startswith("book2"):
        lo, hi = np.array([-1100.0, -50.0, -1100.0]), np.array([1100.0, 600.0, 1100.0])
    n_rand = N_RAYS
    o = rng.uniform(lo, hi, size=(n_rand, 3)).astype(np.float32)
    v = rng.normal(size=(n_rand, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    scale = np.where(rng.random(n_rand) < 0.5, 1.0, rng.uniform(0.05, 2.0, n_rand))
    d = (v * scale[:, None]).astype(np.float32)
    t = rng.random(n_rand).astype(np.float32)
    return o, d, t


def decode_png_rgb(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        ln, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert depth == 8 and ctype == 2
        elif tag == b"IDAT":
            idat += body
        pos += 12 + ln
    raw = zlib.decompress(idat)
    stride = w * 3
    out = np.zeros((h, w, 3), np.uint8)
    prev = np.zeros(stride, np.uint8)
    for y in range(h):
        ft = raw[y * (stride + 1)]
        line = np.frombuffer(raw[y * (stride + 1) + 1:(y + 1) * (stride + 1)], np.uint8).astype(np.int32)
        cur = np.zeros(stride, np.int32)
        for i in range(stride):
            a = cur[i - 3] if i >= 3 else 0
            b = int(prev[i])
            c = int(prev[i - 3]) if i >= 3 else 0
            if ft == 0:
                pred = 0
            elif ft == 1:
                pred = a
            elif ft == 2:
                pred = b
            elif ft == 3:
                pred = (a + b) // 2
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[i] = (line[i] + pred) & 255
        prev = cur.astype(np.uint8)
        out[y] = prev.reshape(w, 3)
    return out


def main():
    for name in SCENES:
        path = os.path.join(ROOT, "data", name + ".json")
        ref = RefScene(path, 16)
        np.save(os.path.join(HERE, f"camera_{name}.npy"), ref.camera())
        np.save(os.path.join(HERE, f"span1_{name}.npy"), ref.span1_flags())
        o, d, t = rays_for(name, 99)
        # Media make Hit() stochastic.  A record is medium-independent iff it is identical in K independent evaluations
        # AND is a surface hit or miss in all of them; rays that ever report a different record are masked out.
        recs = [ref.intersect(o, d, t) for _ in range(12 if ("volume" in name or "book2" in name) else 1)]
        det = np.ones(o.shape[0], bool)
        for r in recs[1:]:
            det &= (r["hit"] == recs[0]["hit"]) & (r["t"].view(np.uint32) == recs[0]["t"].view(np.uint32)) & \
                   (r["material"] == recs[0]["material"])
        r0 = recs[0]
        np.savez_compressed(os.path.join(HERE, f"hits_{name}.npz"), origins=o, directions=d, times=t, hit=r0["hit"], t=r0["t"],
                            point=r0["point"], normal=r0["normal"], front_face=r0["front_face"], material=r0["material"],
                            deterministic=det)
        print(name, "rays", o.shape[0], "hits", int(r0["hit"].sum()), "deterministic", int(det.sum()))
    # low-resolution statistical pins (sum / sumsq in float32 keeps the files small)
    for name, dims, spp in [("cornell_original_test", (60, 60), 256), ("cornell_volume_10000_samples", (60, 60), 256),
                            ("book2_final_scene_10000_samples", (60, 60), 256)]:
        path = os.path.join(ROOT, "data", name + ".json")
        ref = RefScene(path, spp, dims=dims)
        px = None
        if name.startswith("book2"):
            px = ref.perlin_get(0)
        s, ss, rays, sec = ref.render(0, spp, 50, 0, True)
        extra = {}
        if px is not None:
            extra = dict(perm_x=px[0], perm_y=px[1], perm_z=px[2], vec=px[3])
        np.savez_compressed(os.path.join(HERE, f"moments_{name}.npz"), sum=s.astype(np.float32), sumsq=ss.astype(np.float32),
                            spp=spp, dims=np.array(dims), rays=rays, **extra)
        print(name, dims, spp, "rays/path", rays / (dims[0] * dims[1] * spp), f"{sec:.1f}s")
    # WriteImage: fixed float image -> PNG by the reference -> decoded bytes
    rng = np.random.default_rng(5)
    img = rng.uniform(-0.2, 1.3, size=(37, 53, 3)).astype(np.float32)
    img[0, 0] = [0.0, 1.0, 0.25]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "ref.png")
        write_image(img, p, True)
        rgb = decode_png_rgb(p)
    np.savez_compressed(os.path.join(HERE, "tonemap.npz"), image=img, rgb8=rgb)
    print("tonemap", rgb.shape)


if __name__ == "__main__":
    main()
