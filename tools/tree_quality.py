#!/usr/bin/env python3
"""Tree-quality analysis on the CPU (no GPU): node pairs per ray of the extend kernel's walk over (a) the host binned-SAH tree,
(b) the Morton radix tree the device LBVH builder produces and (c) PLOC trees over the same Morton order, for the synthetic sphere scene of BASELINE config 5.

    python tools/tree_quality.py [--n 1000000] [--rays 120000] [--bounces 6] [--ploc 8 32]

Rays: camera rays of a coarse pixel grid, then for every hit a Lambertian bounce (normal + uniform unit vector), repeated —
the mix the renderer traces.  Prints node pairs / sphere tests per ray and bounce for both trees.  Analysis only: nothing here is
on a product or test path (tools/tree_quality.c is compiled on the fly with gcc)."""
import argparse, ctypes as C, os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace2_b200 as rt

HERE = os.path.dirname(os.path.abspath(__file__))
LEAF = 0x80000000


def helper():
    so = os.path.join(tempfile.gettempdir(), "rt2_tree_quality.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "tree_quality.c"), "-lm"])
    lib = C.CDLL(so)
    lib.build_radix_tree.restype = C.c_uint32
    lib.build_ploc.restype = C.c_uint32
    return lib


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def expand21(v):
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x001F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x001F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def radix_tree(lib, bmin, bmax, grid=None):
    """device/rt_lbvh.cu restated: grid = centroid bounds clipped to mean +- 3 sd per axis, 21 bits per axis, x|y|z interleave."""
    n = len(bmin)
    c = (0.5 * (bmin + bmax)).astype(np.float32)
    mean, sd = c.astype(np.float64).mean(0), c.astype(np.float64).std(0)
    lo = np.maximum(c.min(0), (mean - 3 * sd).astype(np.float32))
    hi = np.minimum(c.max(0), (mean + 3 * sd).astype(np.float32))
    if grid is not None:
        lo, hi = np.asarray(grid[0], np.float32), np.asarray(grid[1], np.float32)
    top = np.float32((1 << 21) - 1)
    scale = np.where(hi > lo, top / (hi - lo), 0).astype(np.float32)
    q = np.clip((c - lo) * scale, 0, top).astype(np.uint32)
    keys = (expand21(q[:, 0]) << np.uint64(2)) | (expand21(q[:, 1]) << np.uint64(1)) | expand21(q[:, 2])
    order = np.argsort(keys, kind="stable")
    keys = np.ascontiguousarray(keys[order])
    lmin, lmax = np.ascontiguousarray(bmin[order], np.float32), np.ascontiguousarray(bmax[order], np.float32)
    boxes = np.zeros((n - 1, 2, 6), np.float32)
    entries = np.zeros((n - 1, 2), np.uint32)
    sys.setrecursionlimit(10000)
    pairs = lib.build_radix_tree(ptr(keys, C.c_uint64), ptr(lmin, C.c_float), ptr(lmax, C.c_float), C.c_int(n), ptr(boxes, C.c_float), ptr(entries, C.c_uint32))
    assert pairs == n - 1
    return (boxes, entries, np.ascontiguousarray(order.astype(np.uint32)), None), (lmin, lmax)


def radix_tree_giants_on_top(lib, bmin, bmax, grid=None):
    """The same radix tree over the ordinary primitives only; primitives whose box is > 100 x the median extent (the ground sphere)
    are chained above its root, one pair each."""
    ext = (bmax - bmin).max(1)
    giant = ext > 100.0 * np.median(ext)
    small_idx = np.nonzero(~giant)[0]
    (boxes, entries, order, _), _ = radix_tree(lib, bmin[small_idx], bmax[small_idx], grid)
    prim_ids = list(small_idx[order].astype(np.uint32))
    boxes, entries = list(boxes), list(entries)
    root = 0
    sub_box = np.concatenate([np.minimum(boxes[0][0][:3], boxes[0][1][:3]), np.maximum(boxes[0][0][3:], boxes[0][1][3:])])
    for g in np.nonzero(giant)[0]:
        gbox = np.concatenate([bmin[g], bmax[g]]).astype(np.float32)
        boxes.append(np.stack([gbox, sub_box]).astype(np.float32))
        entries.append(np.array([LEAF | len(prim_ids), root], np.uint32))
        prim_ids.append(np.uint32(g))
        root = len(boxes) - 1
        sub_box = np.concatenate([np.minimum(gbox[:3], sub_box[:3]), np.maximum(gbox[3:], sub_box[3:])])
    return (np.ascontiguousarray(np.array(boxes, np.float32)), np.ascontiguousarray(np.array(entries, np.uint32)),
            np.ascontiguousarray(np.array(prim_ids, np.uint32)), None), root


def ploc_tree(lib, sorted_boxes, order, radius):
    lmin, lmax = sorted_boxes
    n = len(lmin)
    boxes = np.zeros((n - 1, 2, 6), np.float32)
    entries = np.zeros((n - 1, 2), np.uint32)
    root = lib.build_ploc(ptr(lmin, C.c_float), ptr(lmax, C.c_float), C.c_int(n), C.c_int(radius), ptr(boxes, C.c_float), ptr(entries, C.c_uint32))
    return (boxes, entries, order, None), int(root)


def sah_tree(scene):
    nodes = scene.nodes()
    refs = np.ascontiguousarray(scene.prim_refs().astype(np.uint32))
    d = scene.desc
    n_pairs = len(nodes) // 2
    boxes = np.zeros((n_pairs, 2, 6), np.float32)
    boxes[:, :, 0:3] = nodes["bmin"].reshape(n_pairs, 2, 3)
    boxes[:, :, 3:6] = nodes["bmax"].reshape(n_pairs, 2, 3)
    count = nodes["count"].reshape(n_pairs, 2).astype(np.uint32)
    first = nodes["left_first"].reshape(n_pairs, 2).astype(np.uint32)
    entries = np.where(count > 0, LEAF | first, first).astype(np.uint32)
    leaf_count = np.zeros(len(refs) + 1, np.uint32)
    leaf_count[first[count > 0]] = count[count > 0]
    prim_ids = np.ascontiguousarray(refs & np.uint32(0x0FFFFFFF))
    return np.ascontiguousarray(boxes), np.ascontiguousarray(entries), prim_ids, leaf_count, int(d.tlas_root)


def walk(lib, tree, root, spheres, o, d):
    boxes, entries, prim_ids, leaf_count = tree
    n = len(o)
    t = np.zeros(n, np.float32); prim = np.zeros(n, np.int32)
    pairs = np.zeros(n, np.uint32); tests = np.zeros(n, np.uint32); depth = np.zeros(n, np.uint32)
    o, d = np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)
    lib.traverse_count(ptr(boxes, C.c_float), ptr(entries, C.c_uint32), C.c_uint32(root), ptr(prim_ids, C.c_uint32),
                       ptr(leaf_count, C.c_uint32) if leaf_count is not None else None, ptr(spheres, C.c_float), ptr(o, C.c_float),
                       ptr(d, C.c_float), C.c_int(n), C.c_float(0.001), ptr(t, C.c_float), ptr(prim, C.c_int32), ptr(pairs, C.c_uint32),
                       ptr(tests, C.c_uint32), ptr(depth, C.c_uint32))
    return t, prim, pairs, tests, depth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--rays", type=int, default=120000)
    ap.add_argument("--bounces", type=int, default=6)
    ap.add_argument("--ystretch", type=float, nargs="*", default=[], help="also: radix trees with the Morton grid's y range stretched by these factors")
    ap.add_argument("--device-isotropic", action="store_true")
    ap.add_argument("--giants", action="store_true", help="also: radix tree over the ordinary primitives with the giant ones chained above its root")
    ap.add_argument("--ploc", type=int, nargs="*", default=[], help="also build PLOC trees with these search radii")
    a = ap.parse_args()
    lib = helper()
    t0 = time.time()
    scene = rt.Scene.synthetic_spheres(a.n, width=3840, height=2160, host_bvh=True)
    print(f"scene + host SAH build: {time.time() - t0:.1f} s")
    sp = scene.spheres()
    spheres = np.ascontiguousarray(np.concatenate([sp["center0"], sp["radius"][:, None]], axis=1), np.float32)
    bmin, bmax = spheres[:, :3] - spheres[:, 3:4], spheres[:, :3] + spheres[:, 3:4]
    sb, se, sp_ids, slc, sroot = sah_tree(scene)
    t0 = time.time()
    radix, sorted_boxes = radix_tree(lib, bmin, bmax)
    print(f"radix tree: {time.time() - t0:.1f} s")
    trees = {"host binned SAH": ((sb, se, sp_ids, slc), sroot), "Morton radix tree (device LBVH)": (radix, 0)}
    for sy in a.ystretch:
        # Morton grid over the slab's x / z extent, y range stretched by sy (1 = every axis normalised to its own extent,
        # 10 = one isotropic cell size for this 2000 x 200 x 2000 slab)
        g = ([-1000.0, 0.0, -1000.0], [1000.0, 200.0 * sy, 1000.0])
        trees[f"Morton radix tree, y range x {sy:g}"] = (radix_tree(lib, bmin, bmax, g)[0], 0)
        trees[f"Morton radix tree, y range x {sy:g}, giants chained on top"] = radix_tree_giants_on_top(lib, bmin, bmax, g)
    if a.device_isotropic:
        # the isotropic variant as it ran on the GPU: origin = the 3-sigma-clipped centroid minimum, ONE cell size (largest extent)
        c = 0.5 * (bmin + bmax)
        mean, sd = c.astype(np.float64).mean(0), c.astype(np.float64).std(0)
        lo = np.maximum(c.min(0), mean - 3 * sd); hi = np.minimum(c.max(0), mean + 3 * sd)
        span = float((hi - lo).max())
        trees["Morton radix tree, isotropic cells, giant filed in its (clamped) cell"] = (radix_tree(lib, bmin, bmax, (lo, lo + span))[0], 0)
    if a.giants:
        trees["Morton radix tree, giant primitives chained on top"] = radix_tree_giants_on_top(lib, bmin, bmax)
    for r in a.ploc:
        t0 = time.time()
        trees[f"PLOC radius {r} over the same Morton order"] = ploc_tree(lib, sorted_boxes, radix[2], r)
        print(f"PLOC radius {r}: {time.time() - t0:.1f} s")
    # camera rays on a coarse grid
    d = scene.desc; cam = d.camera
    W, H = d.width, d.height
    side = int(np.sqrt(a.rays * W / H))
    xs = (np.arange(side) + 0.5) * W / side
    ys = (np.arange(max(1, side * H // W)) + 0.5) * H / max(1, side * H // W)
    X, Y = np.meshgrid(xs, ys)
    p00, du, dv, c0 = (np.array(list(v), np.float64) for v in (cam.pixel00, cam.pixel_delta_u, cam.pixel_delta_v, cam.center))
    target = p00 + X.reshape(-1, 1) * du + Y.reshape(-1, 1) * dv
    o0 = np.tile(c0, (len(target), 1)); d0 = target - c0
    rng = np.random.default_rng(1)
    res = {}
    for name, (tree, root) in trees.items():
        o, dd = o0.copy(), d0.copy()
        rows = []
        rng = np.random.default_rng(1)
        for b in range(a.bounces):
            if len(o) == 0:
                break
            t, prim, pairs, tests, depth = walk(lib, tree, root, spheres, o, dd)
            rows.append((len(o), pairs.mean(), tests.mean(), int(depth.max())))
            hit = prim >= 0
            if b == 0:
                res.setdefault("first_hits", {})[name] = (prim.copy(), t.copy())
            p = o[hit] + dd[hit] * t[hit, None].astype(np.float64)
            nrm = (p - spheres[prim[hit], :3]) / spheres[prim[hit], 3:4]
            front = np.sum(nrm * dd[hit], axis=1) < 0
            nrm = np.where(front[:, None], nrm, -nrm)
            u = rng.normal(size=nrm.shape); u /= np.linalg.norm(u, axis=1, keepdims=True)
            o, dd = p, nrm + u
        res[name] = rows
    firsts = list(res["first_hits"].values())
    print("closest hits of all trees agree:", all(bool(np.array_equal(firsts[0][0], f[0])) for f in firsts[1:]))
    for name in trees:
        tot_r = sum(r[0] for r in res[name]); tot_p = sum(r[0] * r[1] for r in res[name]); tot_t = sum(r[0] * r[2] for r in res[name])
        print(f"\n{name}: {tot_p / tot_r:.1f} node pairs / ray, {tot_t / tot_r:.2f} sphere tests / ray over {tot_r} rays")
        for b, r in enumerate(res[name]):
            print(f"   bounce {b}: {r[0]:7d} rays  {r[1]:6.1f} pairs  {r[2]:5.2f} tests  max stack {r[3]}")


if __name__ == "__main__":
    main()
