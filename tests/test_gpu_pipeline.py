"""GPU: properties of the wavefront pipeline that do not need the oracle.  Philox is keyed on (pixel, frame, bounce), never
on queue positions, so: sorted and unsorted queues render the SAME image bit for bit; the fused and the per-bin shade
pipelines trace the same paths (identical except where the compiler contracted the shading arithmetic differently in the
two kernels — a last-bit change of a direction that later flips a hit); the flat and the BVH extend kernels report the
same hits; and an interrupted + resumed render equals an uninterrupted one."""
import os

import numpy as np
import pytest

import raytrace2_b200 as rt
from conftest import scene_path
from _rays import fixed_rays

pytestmark = pytest.mark.gpu


def _render(name, dims, spp, flags=0, fpb=0, seed=77):
    scene = rt.Scene.load(scene_path(name), perlin_seed=5)
    tr = rt.RayTracer(scene, num_samples=spp, max_depth=50, seed=seed, flags=flags, dims=dims, frames_per_batch=fpb)
    tr.Update(spp)
    return tr, tr.read_accum()


@pytest.mark.parametrize("name,dims,spp", [("book2_final_scene_10000_samples", (200, 200), 16), ("cornell_volume_10000_samples", (160, 160), 16),
                                           ("final_render_book_1", (320, 180), 9), ("checker_test", (160, 90), 16)])
def test_fused_and_per_bin_pipelines_trace_the_same_paths(native_lib, name, dims, spp):
    ta, a = _render(name, dims, spp)
    tb, b = _render(name, dims, spp, flags=rt.RT2_FLAG_NO_FUSED_SHADE)
    ra, rb = ta.stats()["rays"], tb.stats()["rays"]
    assert abs(ra - rb) < 1e-3 * ra, (ra, rb)
    same = np.all(a.view(np.uint32) == b.view(np.uint32), axis=-1)
    assert same.mean() > 0.97, same.mean()  # shading is plain (FMA-contractable) float arithmetic: parity there is statistical
    assert abs(float(a.mean()) - float(b.mean())) < 0.02 * float(a.mean()) + 1e-6
    assert ta.stats()["launches"] < tb.stats()["launches"]


def test_sorted_queue_renders_the_same_image(native_lib):
    """The ray sort measured as a net loss (profiles/r01_notes.md) and lives only in `make EXPERIMENTS=1` builds: the shipped
    library must reject the flag loudly; an experiments build must render the same image with and without it."""
    name, dims, spp = "book2_final_scene_10000_samples", (256, 256), 16
    _, a = _render(name, dims, spp)
    try:
        tb, b = _render(name, dims, spp, flags=rt.RT2_FLAG_SORT_RAYS)
    except rt.Rt2Error as e:
        assert e.code == -5 and "EXPERIMENTS" in e.message
        return
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", ["cornell_original_test", "cornell_box_scene_graph", "cornell_volume_10000_samples"])
def test_flat_and_bvh_extend_agree(native_lib, name):
    """Tiny scenes take k_traverse_flat; the BVH walk (RT2_FLAG_NO_FLAT_EXTEND) must report the same t bit for bit and the same primitive
    except on exact ties between coincident faces (resolved by test order)."""
    scene = rt.Scene.load(scene_path(name))
    o, d, _ = fixed_rays(scene, 60000, seed=9)
    tf = rt.RayTracer(scene, dims=(64, 64))
    tb = rt.RayTracer(scene, dims=(64, 64), flags=rt.RT2_FLAG_NO_FLAT_EXTEND)  # the BVH walk over the unified world tree
    n_inst = len(scene.instances())  # the Cornell boxes are instances; box-bounded media under a transform are not
    assert tb.stats()["instance_mode"] == (3 if n_inst else 0) and tf.stats()["instance_mode"] == 4
    f = tf.intersect(o, d, skip_media=True)
    b = tb.intersect(o, d, skip_media=True)
    assert np.array_equal(f["material"] >= 0, b["material"] >= 0)
    hit = f["material"] >= 0
    assert hit.sum() > 10000
    assert np.array_equal(f["t"][hit].view(np.uint32), b["t"][hit].view(np.uint32))
    same = (f["prim"][hit] == b["prim"][hit]) & (f["instance"][hit] == b["instance"][hit])
    assert same.mean() > 0.999  # ties follow the reference's rule in both kernels (quad_wins_tie)
    # the render kernels: statistically the same image is checked against the oracle elsewhere; here the ray counts
    tf.Update(4)
    tb.Update(4)
    ra, rb = tf.stats()["rays"], tb.stats()["rays"]
    assert abs(ra - rb) < 0.02 * ra


def test_checkpoint_resume_is_bit_identical(native_lib, tmp_path):
    name, dims = "book2_final_scene_10000_samples", (128, 128)
    scene = rt.Scene.load(scene_path(name), perlin_seed=5)
    kw = dict(num_samples=64, max_depth=50, seed=123, flags=rt.RT2_FLAG_MOMENTS, dims=dims, frames_per_batch=8)
    straight = rt.RayTracer(scene, **kw)
    straight.Update(24)
    s0, ss0 = straight.read_accum(moments=True)
    first = rt.RayTracer(scene, **kw)
    first.Update(8)
    ck = str(tmp_path / "ck.npz")
    first.save_checkpoint(ck)
    del first
    second = rt.RayTracer(scene, **kw)
    assert second.load_checkpoint(ck) == 8 and second.FrameIdx() == 8
    second.Update(16)
    s1, ss1 = second.read_accum(moments=True)
    assert second.FrameIdx() == 24
    assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32)) and np.array_equal(ss0.view(np.uint32), ss1.view(np.uint32))
    assert np.array_equal(straight.NonConvertedPixels().view(np.uint32), second.NonConvertedPixels().view(np.uint32))
    other = rt.RayTracer(scene, **{**kw, "seed": 124})
    with pytest.raises(ValueError):
        other.load_checkpoint(ck)


def test_resolve_peers_with_one_rank_equals_the_plain_read_out(native_lib):
    """rt2_resolve_peers on a single rank (no peer mapping needed) is NonConvertedPixels / Pixels; the N-rank path runs under
    torchrun in tools/dist_check.py (CUDA IPC cannot map an allocation into the process that owns it)."""
    scene = rt.Scene.load(scene_path("cornell_original_test"))
    tr = rt.RayTracer(scene, num_samples=16, seed=4, dims=(96, 96))
    tr.Update(16)
    handle = tr.accum_ipc_handle()
    assert len(handle) == 80 and any(handle[:64])
    a = tr.resolve_peers([handle], 0, 16)
    assert np.array_equal(a.view(np.uint32), tr.NonConvertedPixels().view(np.uint32))
    assert np.array_equal(tr.resolve_peers([handle], 0, 16, rgba8=True), tr.Pixels())
    with pytest.raises(rt.Rt2Error):
        tr.resolve_peers([handle], 3, 16)


def test_two_gpu_read_out_paths(native_lib):
    """tools/dist_check.py under torchrun with 2 ranks: peer-memory read-out == NCCL read-out == single-GPU render.  Needs two
    visible GPUs (skipped on a one-GPU box; the driver's multi-GPU runs and profiles/r01_notes.md cover it there)."""
    import subprocess
    import sys
    from conftest import ROOT
    if native_lib.rt2_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29633", os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dist_check OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
