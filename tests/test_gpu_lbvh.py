"""GPU LBVH builder (raytrace2_b200/csrc/device/rt_lbvh.cu, RT2_FLAG_GPU_LBVH): structural invariants of the device-built
tree and independence of the closest-hit results from the tree (result = arg-min over leaves of raw t)."""
import numpy as np
import pytest

import raytrace2_b200 as rt
from _rays import fixed_rays
from conftest import scene_path

pytestmark = pytest.mark.gpu


def check_tree(nodes, refs, roots, n_skip_refs, expect_refs):
    """Every reference of [n_skip_refs, len(refs)) is reached exactly once from the given roots; child boxes lie inside
    their parent's box; leaves hold one primitive; depth fits the 64-entry device stack."""
    seen = np.zeros(len(refs), np.int32)
    max_depth = 0
    for root in roots:
        stack = [(int(root), None, None, 0)]
        while stack:
            pair, lo, hi, depth = stack.pop()
            max_depth = max(max_depth, depth)
            for side in range(2):
                n = nodes[2 * pair + side]
                bmin, bmax = np.array(n["bmin"]), np.array(n["bmax"])
                if not (bmin[0] <= bmax[0]):
                    continue
                if lo is not None:
                    assert np.all(bmin >= lo - 1e-3 * (1 + np.abs(lo))) and np.all(bmax <= hi + 1e-3 * (1 + np.abs(hi)))
                if n["count"] == 0:
                    stack.append((int(n["left_first"]), bmin, bmax, depth + 1))
                else:
                    assert n["count"] == 1
                    seen[int(n["left_first"])] += 1
    assert max_depth < 60, max_depth
    assert (seen[:n_skip_refs] == 0).all() and (seen[n_skip_refs:] == 1).all()
    assert sorted(refs[n_skip_refs:].tolist()) == sorted(expect_refs.tolist())
    return max_depth


def _host_tree_refs(scene):
    """Primitive references reachable from the host SAH trees (everything except the media boundary lists)."""
    refs = scene.prim_refs()
    skip = set()
    for m in scene.media():
        skip.update(range(int(m["boundary_first"]), int(m["boundary_first"] + m["boundary_count"])))
    return np.array([r for i, r in enumerate(refs) if i not in skip], np.uint32)


@pytest.mark.parametrize("builder", [0, rt.RT2_FLAG_LBVH_PLOC])
@pytest.mark.parametrize("name", ["book2_final_scene_10000_samples", "cornell_original_test", "cornell_box_scene_graph",
                                  "cornell_volume_10000_samples", "final_render_book_1"])
def test_device_built_tree_is_valid_and_gives_identical_hits(native_lib, name, builder):
    """builder: Karras' radix tree over the 63-bit Morton order (default) or PLOC clustering (opt-in; measured worse trees)."""
    scene = rt.Scene.load(scene_path(name))
    sah = rt.RayTracer(scene)
    lbvh = rt.RayTracer(scene, flags=rt.RT2_FLAG_GPU_LBVH | builder)
    nodes, refs, root = lbvh.read_bvh()
    n_media_refs = int(sum(m["boundary_count"] for m in scene.media()))
    # tree 0 (TLAS) starts at pair 0; instance BLAS roots are patched into the device copy of the instance table, so walk
    # the TLAS only and count instance leaves, then walk every BLAS root we can infer from the layout: pairs are laid out
    # tree after tree, each tree of n prims owning max(1, n-1) pairs.
    d = scene.desc
    host_refs = _host_tree_refs(scene)
    # recover the per-tree primitive counts from the host instance table (BLAS sizes) — same order as the device layout
    host_nodes = scene.nodes()
    host_prim_refs = scene.prim_refs()

    def count_leaves(root_pair):
        total, st = 0, [int(root_pair)]
        while st:
            p = st.pop()
            for side in range(2):
                nd = host_nodes[2 * p + side]
                if not (nd["bmin"][0] <= nd["bmax"][0]):
                    continue
                if nd["count"] == 0:
                    st.append(int(nd["left_first"]))
                else:
                    total += int(nd["count"])
        return total
    sizes = [count_leaves(d.tlas_root)] + [count_leaves(int(i["blas_root"])) for i in scene.instances()]
    if d.has_world_tlas:  # instance split: one more world tree over the surfaces only, laid out last
        sizes.append(count_leaves(d.tlas_world_root))
    if d.has_unified_tlas:  # ... and the unified world tree (surfaces + instanced primitives as world-space leaves)
        sizes.append(count_leaves(d.tlas_unified_root))
    roots, base = [], 0
    for n in sizes:
        roots.append(base)
        base += max(1, n - 1)
    assert base == len(nodes) // 2
    check_tree(nodes, refs, roots, n_media_refs, host_refs)
    assert root == 0
    # identical closest hits from both trees
    o, dd, tm = fixed_rays(scene, 100000, seed=17)
    a = sah.intersect(o, dd, tm, skip_media=True)
    b = lbvh.intersect(o, dd, tm, skip_media=True)
    assert np.array_equal(a["material"] >= 0, b["material"] >= 0)
    same = (a["t"].view(np.uint32) == b["t"].view(np.uint32)) & (a["prim"] == b["prim"])
    hit = a["material"] >= 0
    # exact ties between coincident faces may resolve differently (traversal order); everything else is bit-identical
    assert (hit & ~same).sum() <= 0.03 * hit.sum()
    tie = hit & ~same
    assert np.array_equal(a["t"][tie].view(np.uint32), b["t"][tie].view(np.uint32)), "only ties may differ between trees"
    ok = hit & same
    assert np.array_equal(a["point"][ok].view(np.uint32), b["point"][ok].view(np.uint32))
    assert np.array_equal(a["normal"][ok], b["normal"][ok]) and np.array_equal(a["instance"][ok], b["instance"][ok])


def test_synthetic_scene_without_host_bvh(native_lib):
    """BASELINE config 5 path: the scene skips the host SAH build and can only be rendered with the device-built tree."""
    n = 200000
    with_host = rt.Scene.synthetic_spheres(n, seed=9, width=320, height=180, host_bvh=True)
    no_host = rt.Scene.synthetic_spheres(n, seed=9, width=320, height=180, host_bvh=False)
    assert no_host.desc.n_node_pairs == 0 and no_host.desc.n_spheres == n + 1
    with pytest.raises(rt.Rt2Error) as e:
        rt.RayTracer(no_host)
    assert e.value.code == -6 and "RT2_FLAG_GPU_LBVH" in e.value.message
    sah = rt.RayTracer(with_host, num_samples=16, seed=2)
    lbvh = rt.RayTracer(no_host, num_samples=16, seed=2, flags=rt.RT2_FLAG_GPU_LBVH)
    st = lbvh.stats()
    assert 0.0 < st["gpu_ms_bvh_build"] < 200.0
    nodes, refs, root = lbvh.read_bvh()
    assert len(nodes) // 2 == n and len(refs) == n + 1
    depth = check_tree(nodes, refs, [0], 0, with_host.prim_refs())
    assert depth >= 17
    o, dd, tm = fixed_rays(with_host, 200000, seed=23)
    a, b = sah.intersect(o, dd, tm), lbvh.intersect(o, dd, tm)
    # Tree independence holds up to the float false positives of Sphere::Hit on distant small spheres: such a
    # "hit" lies outside the sphere's own box, so whether it is ever evaluated depends on the leaf boxes (SAH leaves hold up
    # to 4 spheres, LBVH leaves one).  It only affects rays that start ~3000 units away (the camera rays, a quarter of the
    # set): at this sphere density ~0.6 % of them.  Bar: <= 0.3 % of all rays, none among the interior rays' near hits, and in
    # every disagreement one side is geometrically outside its sphere.
    diff = np.nonzero((a["t"].view(np.uint32) != b["t"].view(np.uint32)) | (a["prim"] != b["prim"]))[0]
    assert len(diff) <= 3e-3 * len(o), len(diff)
    assert (np.minimum(a["t"][diff], b["t"][diff]) > 300.0).all(), "near hits must not depend on the tree"
    sph = with_host.spheres()

    def outside(rec):
        if rec["material"] < 0:
            return False
        s = sph[int(rec["prim"]) & 0x0FFFFFFF]
        dist = np.linalg.norm(rec["point"].astype(np.float64) - np.array(s["center0"], np.float64))
        return dist > float(s["radius"]) * 1.002  # a true hit point lies ON the sphere (|p - c| = r up to ~3e-4 relative)
    for i in diff:
        assert outside(a[i]) or outside(b[i]), f"ray {i}: trees disagree on a genuine hit"
    # same rays, same RNG keys, (almost) the same hits => most pixels are bit-identical (a path that meets one of the
    # noise hits above diverges afterwards: ~1 % of samples, so up to ~15 % of the 16-spp pixels) and the means agree
    sah.Update(16)
    lbvh.Update(16)
    ia, ib = sah.read_accum(), lbvh.read_accum()
    frac_px = (ia != ib).any(axis=2).mean()
    ra, rb = sah.stats()["rays"], lbvh.stats()["rays"]
    assert frac_px < 0.2, frac_px
    assert abs(ia.mean() - ib.mean()) < 1e-2 * ia.mean(), (ia.mean(), ib.mean())
    assert abs(ra - rb) < 1e-2 * ra, (ra, rb)


def test_radix_sort_orders_morton_codes(native_lib):
    """The tree's leaf order is the sorted Morton order: consecutive leaves are spatially close (sanity of sort + hierarchy)."""
    scene = rt.Scene.synthetic_spheres(50000, seed=4, width=64, height=64, host_bvh=False)
    tr = rt.RayTracer(scene, flags=rt.RT2_FLAG_GPU_LBVH)
    nodes, refs, _ = tr.read_bvh()
    sph = scene.spheres()
    c = np.array([sph[int(r) & 0x0FFFFFFF]["center0"] for r in refs], np.float64)
    small = np.array([sph[int(r) & 0x0FFFFFFF]["radius"] < 10 for r in refs])
    cc = c[small]
    step = np.linalg.norm(np.diff(cc, axis=0), axis=1)
    rnd = np.linalg.norm(cc[np.random.default_rng(0).permutation(len(cc))][1:] - cc[:-1], axis=1)
    assert np.median(step) < 0.1 * np.median(rnd)


# ---- 4-wide quantised tree (RT2_FLAG_WIDE_BVH, rt_wide.cuh) ------------------------------------------------------------
def _wide_vs_binary(scene, n_rays, seed):
    o, d, t = fixed_rays(scene, n_rays, seed=seed)
    binary = rt.RayTracer(scene, flags=rt.RT2_FLAG_GPU_LBVH, dims=(64, 64))
    wide = rt.RayTracer(scene, flags=rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_WIDE_BVH, dims=(64, 64))
    b = binary.intersect(o, d, t, skip_media=True)
    w = wide.intersect(o, d, t, skip_media=True)
    assert np.array_equal(b["material"] >= 0, w["material"] >= 0)
    hit = b["material"] >= 0
    assert hit.sum() > n_rays // 20
    # Same closest t, bit for bit — except where the reference arithmetic itself depends on WHICH spheres are tested: for a
    # small sphere thousands of units from the ray origin the discriminant of Sphere::Hit cancels to an absolute error of ~1,
    # so a sphere whose (looser, quantised) box the ray crosses can report a hit the other tree never evaluates.
    same_t = b["t"][hit].view(np.uint32) == w["t"][hit].view(np.uint32)
    assert same_t.mean() > 0.998, same_t.mean()
    assert np.array_equal(b["point"][hit][same_t].view(np.uint32), w["point"][hit][same_t].view(np.uint32))
    assert (b["prim"][hit][same_t] == w["prim"][hit][same_t]).mean() > 0.999                # exact ties aside
    return binary, wide


def test_wide_tree_gives_identical_hits_on_the_sphere_field(native_lib):
    scene = rt.Scene.synthetic_spheres(200000, width=256, height=144, host_bvh=False)
    binary, wide = _wide_vs_binary(scene, 80000, seed=12)
    # the renders trace the same paths: same ray totals, same image
    binary.Update(4)
    wide.Update(4)
    ra, rw = binary.stats()["rays"], wide.stats()["rays"]
    assert abs(ra - rw) < 2e-3 * ra
    same = np.all(binary.read_accum().view(np.uint32) == wide.read_accum().view(np.uint32), axis=-1)
    assert same.mean() > 0.98


def test_wide_tree_on_a_scene_file_and_its_limits(native_lib):
    scene = rt.Scene.load(scene_path("final_render_book_1"))   # 484 spheres incl. the r = 1000 ground sphere, no instances
    _wide_vs_binary(scene, 60000, seed=13)
    with pytest.raises(rt.Rt2Error):                            # instances are not supported by the single wide tree
        rt.RayTracer(rt.Scene.load(scene_path("cornell_original_test")), flags=rt.RT2_FLAG_GPU_LBVH | rt.RT2_FLAG_WIDE_BVH)
    with pytest.raises(rt.Rt2Error):                            # needs the device-built tree
        rt.RayTracer(scene, flags=rt.RT2_FLAG_WIDE_BVH)
