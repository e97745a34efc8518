#!/usr/bin/env python3
"""First-light GPU check: rt2_intersect vs the compiled reference on fixed rays, then a short render vs the reference.
Run under gpurun from the repo root:  python tools/gpu_check.py [--spp 16]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt  # noqa: E402
from raytrace2_b200 import parity  # noqa: E402
from oracle.ref_oracle import RefScene  # noqa: E402


def make_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    nodes = scene.nodes()
    d = scene.desc
    root = nodes[2 * d.tlas_root: 2 * d.tlas_root + 2]
    lo = np.minimum(*[np.array(x["bmin"]) for x in root if x["bmin"][0] <= x["bmax"][0]]) if len(root) else np.zeros(3)
    hi = np.maximum(*[np.array(x["bmax"]) for x in root if x["bmin"][0] <= x["bmax"][0]]) if len(root) else np.ones(3)
    lo, hi = np.maximum(lo, -2000), np.minimum(hi, 2000)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    scale = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(0.05, 2.0, n))
    dirs = (v * scale[:, None]).astype(np.float32)
    t = rng.random(n).astype(np.float32)
    return o, dirs, t


def check_intersect(path, n=200000):
    scene = rt.Scene.load(path, data_dir="data")
    ref = RefScene(path, 16)
    tracer = rt.RayTracer(scene, num_samples=16)
    o, d, tm = make_rays(scene, n, 1234)
    t0 = time.time()
    g = tracer.intersect(o, d, tm, skip_media=True)
    t1 = time.time()
    r = ref.intersect(o, d, tm)
    t2 = time.time()
    mats = scene.materials()
    iso = np.array([m["type"] == 5 for m in mats])
    ref_hit = r["hit"].astype(bool)
    ref_medium = ref_hit & iso[np.clip(r["material"], 0, len(mats) - 1)]
    cmp_mask = ~ref_medium
    g_hit = g["material"] >= 0
    both = cmp_mask & ref_hit & g_hit
    hit_mismatch = int((cmp_mask & (ref_hit != g_hit)).sum())
    t_bits = int((g["t"][both].view(np.uint32) != r["t"][both].view(np.uint32)).sum())
    p_bits = int((g["point"][both].view(np.uint32) != r["point"][both].view(np.uint32)).any(axis=1).sum())
    n_bits = int((g["normal"][both].view(np.uint32) != r["normal"][both].view(np.uint32)).any(axis=1).sum())
    m_diff = int((g["material"][both] != r["material"][both]).sum())
    ff_diff = int((g["front_face"][both] != r["front_face"][both]).sum())
    rel = np.abs(g["t"][both] - r["t"][both]) / np.maximum(np.abs(r["t"][both]), 1e-30)
    print(f"[intersect] {os.path.basename(path)}: rays={n} hits={int(both.sum())} medium_excluded={int(ref_medium.sum())} "
          f"hit_flag_mismatch={hit_mismatch} t_bit_diff={t_bits} point_bit_diff={p_bits} normal_bit_diff={n_bits} "
          f"material_diff={m_diff} front_face_diff={ff_diff} max_rel_t={float(rel.max()) if rel.size else 0:.3e} "
          f"gpu_s={t1 - t0:.3f} ref_s={t2 - t1:.3f}")
    return hit_mismatch, t_bits, m_diff


def check_render(path, spp, ref_spp):
    scene = rt.Scene.load(path, data_dir="data")
    ref = RefScene(path, spp)
    if scene.desc.n_perlin:
        px, py, pz, vec = scene.get_perlin(0)
        for ti, t in enumerate(scene.textures()):
            if t["type"] == 2:
                ref.perlin_set(ti, px, py, pz, vec)
    tracer = rt.RayTracer(scene, num_samples=spp, flags=rt.RT2_FLAG_MOMENTS)
    t0 = time.time()
    tracer.Update(spp)
    tracer.synchronize()
    t1 = time.time()
    st = tracer.stats()
    s, ss = tracer.read_accum(moments=True)
    rs, rss, rrays, rsec = ref.render(0, ref_spp, 50, 0, True)
    z, valid = parity.z_scores(s, ss, spp, rs, rss, ref_spp)
    tz = parity.tile_z_scores(s, ss, spp, rs, rss, ref_spp, 50)
    w, h = tracer.Dims()
    print(f"[render] {os.path.basename(path)}: gpu spp={spp} wall={t1 - t0:.3f}s gpu_ms={st['gpu_ms_total']:.1f} rays={st['rays']} "
          f"rays/path={st['rays'] / max(st['paths'], 1):.3f} Mrays/s={st['rays'] / max(st['gpu_ms_total'], 1e-9) * 1e-3:.1f} | "
          f"ref spp={ref_spp} s={rsec:.2f} rays/path={rrays / (w * h * ref_spp):.3f} Mrays/s={rrays / rsec * 1e-6:.2f}")
    print(f"         mean gpu={s.mean() / spp:.5f} ref={rs.mean() / ref_spp:.5f}  z: {parity.summary(z, valid)}  "
          f"tile|z| max={np.abs(tz).max():.2f} mean={np.abs(tz).mean():.2f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--ref-spp", type=int, default=16)
    ap.add_argument("--scenes", nargs="*", default=["cornell_original_test", "cornell_box_scene_graph", "cornell_box4",
                                                     "cornell_volume_10000_samples", "book2_final_scene_10000_samples"])
    a = ap.parse_args()
    print("devices:", rt.load_library().rt2_device_count())
    for s in a.scenes:
        check_intersect(f"data/{s}.json")
    for s in a.scenes:
        check_render(f"data/{s}.json", a.spp, a.ref_spp)


if __name__ == "__main__":
    main()
