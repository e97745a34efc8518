"""Deterministic fixed-ray sets for the intersection parity tests: camera rays through pixel centres
plus rays with random origins inside the scene bounds, half unit-length and half |d| in (0.05, 2] (exercises quirk Q1)."""
import numpy as np


def scene_bounds(scene):
    d = scene.desc
    nodes = scene.nodes()
    root = nodes[2 * d.tlas_root: 2 * d.tlas_root + 2]
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for n in root:
        if n["bmin"][0] <= n["bmax"][0]:  # skips empty slots (NaN bounds)
            lo = np.minimum(lo, np.array(n["bmin"]))
            hi = np.maximum(hi, np.array(n["bmax"]))
    lo, hi = np.maximum(lo, -1500.0), np.minimum(hi, 1500.0)
    bad = ~(np.isfinite(lo) & np.isfinite(hi) & (lo < hi))  # e.g. a scene without a host tree: sample a generic box
    lo[bad], hi[bad] = -1500.0, 1500.0
    return lo, hi


def fixed_rays(scene, n: int, seed: int):
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(scene)
    d = scene.desc
    n_cam = n // 4
    n_rand = n - n_cam
    # camera rays: pixel centres, from the camera centre (no lens, no jitter)
    cam = d.camera
    xs = rng.integers(0, d.width, n_cam)
    ys = rng.integers(0, d.height, n_cam)
    p00, du, dv, c = (np.array(list(v), np.float32) for v in (cam.pixel00, cam.pixel_delta_u, cam.pixel_delta_v, cam.center))
    pc = p00[None] + xs[:, None].astype(np.float32) * du[None] + ys[:, None].astype(np.float32) * dv[None]
    dc = pc - c[None]
    dc = (dc / np.linalg.norm(dc, axis=1, keepdims=True)).astype(np.float32)
    oc = np.repeat(c[None], n_cam, 0)
    # interior rays
    o = rng.uniform(lo, hi, size=(n_rand, 3)).astype(np.float32)
    v = rng.normal(size=(n_rand, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    scale = np.where(rng.random(n_rand) < 0.5, 1.0, rng.uniform(0.05, 2.0, n_rand))
    dr = (v * scale[:, None]).astype(np.float32)
    origins = np.concatenate([oc, o]).astype(np.float32)
    dirs = np.concatenate([dc, dr]).astype(np.float32)
    times = rng.random(n).astype(np.float32)
    return origins, dirs, times


def leaf_to_prim_ref(scene_json_path):
    """Maps the CPU restatement's leaf ids (creation order of spheres / quads / box faces / medium wrappers, see
    oracle/rt_oracle.cpp) to the product's primitive references ((type << 28) | index; include/rt2.h)."""
    import json
    doc = json.load(open(scene_json_path))
    prims = doc["primitives"]
    out = []
    n_sph = n_quad = n_med = 0
    if isinstance(prims, dict):  # legacy: spheres, quads, boxes in that order
        seq = [("sphere", p) for p in prims.get("spheres", [])] + [("quad", p) for p in prims.get("quads", [])] + \
              [("box", p) for p in prims.get("boxes", [])]
    else:
        seq = [(p.get("type", ""), p) for p in prims]
    for ty, p in seq:
        if ty == "sphere":
            out.append((0 << 28) | n_sph)
            n_sph += 1
        elif ty == "quad":
            out.append((1 << 28) | n_quad)
            n_quad += 1
        elif ty == "box":
            for _ in range(6):
                out.append((1 << 28) | n_quad)
                n_quad += 1
        else:
            continue
        cm = p.get("constant_medium")
        if cm is not None and ("albedo" in cm or "material" in cm):
            out.append((3 << 28) | n_med)
            n_med += 1
    return np.array(out + [0xFFFFFFFF], np.uint32)  # index -1 (miss) -> PRIM_NONE
