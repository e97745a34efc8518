"""Arithmetic of the 32-byte quantised node pairs (raytrace2_b200/csrc/device/rt_qnodes.cu, rt_trace.cuh), restated in numpy:

  * grid set-up (QuantiseNodesOnDevice): origin 16 cells below the scene's min corner, extent = span * (1 + 1/512);
  * encode (k_quantise_nodes): q_lo = floor((lo - origin) * 32768 / ext) - 1, q_hi = ceil(...) + 1, stored as 0x8000 | q;
  * decode (traverse_queue, kQuant): PRMT(word, 0x3F000000, selector) = bytes {00, lo, hi, 3F} = the float 1 + q / 32768, and the
    plane is base + v * ext with base = origin - ext;
  * the per-ray selector (sign of the direction) picks the entry plane from the low half (min) or the high half (max) of a word.

The property the GPU parity tests rely on (tests/test_gpu_round2.py::test_compact_nodes_*): the decoded box CONTAINS the float box
with at least half a cell to spare, so box tests — which only cull — can never lose a leaf.  No GPU needed: this pins the formulas."""
import numpy as np

F = np.float32


def grid_for(lo, hi):
    lo, hi = np.asarray(lo, F), np.asarray(hi, F)
    span_max = F(np.max(hi - lo))
    span = np.maximum(hi - lo, span_max * F(1e-6)).astype(F)
    ext = (span * F(1.0 + 1.0 / 512.0)).astype(F)
    origin = (lo - span * F(1.0 / 2048.0)).astype(F)
    inv_cell = (F(32768.0) / ext).astype(F)
    return origin, ext, inv_cell, (origin - ext).astype(F)


def encode(lo, hi, origin, inv_cell):
    a = np.floor(((lo - origin).astype(F) * inv_cell).astype(F)) - F(1.0)
    b = np.ceil(((hi - origin).astype(F) * inv_cell).astype(F)) + F(1.0)
    clamped = ~((a >= 0) & (b <= 32767))
    qlo = np.clip(a, 0, 32767).astype(np.uint32)
    qhi = np.clip(b, 0, 32767).astype(np.uint32)
    return (0x8000 | qlo) | ((0x8000 | qhi) << 16), clamped  # one word per axis: min | max << 16


def prmt(a, b, sel):
    """PTX prmt.b32 (default mode) for selectors whose nibbles are 0..7."""
    src = np.concatenate([np.asarray(a, np.uint32).reshape(-1, 1).view(np.uint8), np.asarray(b, np.uint32).reshape(-1, 1).view(np.uint8)], axis=1)
    idx = [(sel >> (4 * k)) & 0xF for k in range(4)]
    out = np.stack([src[:, i] for i in idx], axis=1).copy()
    return out.view(np.uint32).reshape(-1)


SEL_LOW, SEL_HIGH = 0x7104, 0x7324


def decode(words, sel):
    return prmt(words, np.full(len(words), 0x3F000000, np.uint32), sel).view(F)


def test_prmt_selector_makes_one_plus_q_over_32768():
    q = np.array([0, 1, 2, 255, 256, 4097, 32766, 32767], np.uint32)
    words = (0x8000 | q) | ((0x8000 | q[::-1]) << 16)
    assert np.array_equal(decode(words, SEL_LOW), (1.0 + q / 32768.0).astype(F))
    assert np.array_equal(decode(words, SEL_HIGH), (1.0 + q[::-1] / 32768.0).astype(F))
    assert SEL_LOW ^ 0x0220 == SEL_HIGH  # the far-plane selector is the near-plane one with the two source nibbles flipped


def test_decoded_box_contains_float_box():
    rng = np.random.default_rng(5)
    for trial in range(20):
        scale = 10.0 ** rng.uniform(-1, 4)
        centre = rng.uniform(-1, 1, 3) * scale * rng.choice([0.0, 1.0, 50.0])
        scene_lo, scene_hi = (centre - scale).astype(F), (centre + scale * rng.uniform(0.2, 1.0, 3)).astype(F)
        origin, ext, inv_cell, base = grid_for(scene_lo, scene_hi)
        n = 4000
        u = rng.uniform(0, 1, (n, 3))
        w = rng.uniform(0, 1, (n, 3)) ** 4
        lo = (scene_lo + (scene_hi - scene_lo) * u * (1 - w)).astype(F)
        hi = np.minimum(lo + ((scene_hi - scene_lo) * w).astype(F), scene_hi).astype(F)
        lo[0], hi[0] = scene_lo, scene_hi  # the root box itself
        hi[1] = lo[1]                      # a degenerate (flat) box
        for k in range(3):
            words, clamped = encode(lo[:, k], hi[:, k], origin[k], inv_cell[k])
            assert not clamped.any(), "boxes inside the scene's bounds never reach the clamp"
            cell = float(ext[k]) / 32768.0
            x_min = base[k].astype(np.float64) + decode(words, SEL_LOW).astype(np.float64) * float(ext[k])
            x_max = base[k].astype(np.float64) + decode(words, SEL_HIGH).astype(np.float64) * float(ext[k])
            assert np.all(x_min <= lo[:, k] - 0.5 * cell), (trial, k)
            assert np.all(x_max >= hi[:, k] + 0.5 * cell), (trial, k)
            assert np.all(x_min >= lo[:, k] - 2.5 * cell) and np.all(x_max <= hi[:, k] + 2.5 * cell), "at most ~2 cells of padding"


def test_folded_slab_matches_plain_slab_within_the_padding():
    """t = v * (ext / d) + (base - o) / d, as the kernel evaluates it in binary32, against the plane's parameter in binary64:
    the error is a small fraction of one cell's worth of t, i.e. far inside the one-cell padding."""
    rng = np.random.default_rng(9)
    scene_lo, scene_hi = np.array([-1000, 0, -1000], F), np.array([1000, 560, 1000], F)
    origin, ext, inv_cell, base = grid_for(scene_lo, scene_hi)
    n = 20000
    o = rng.uniform(scene_lo, scene_hi, (n, 3)).astype(F)
    d = rng.normal(size=(n, 3)).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(F)
    q = rng.integers(0, 32768, (n, 3)).astype(np.uint32)
    for k in range(3):
        v = decode((0x8000 | q[:, k]).astype(np.uint32), SEL_LOW)
        inv = (F(1.0) / d[:, k]).astype(F)
        oid = ((base[k] - o[:, k]).astype(F) * inv).astype(F)
        inv_q = (ext[k] * inv).astype(F)
        t32 = (v.astype(np.float64) * inv_q.astype(np.float64) + oid.astype(np.float64)).astype(F)  # one fused multiply-add
        plane = float(base[k]) + v.astype(np.float64) * float(ext[k])
        t64 = (plane - o[:, k].astype(np.float64)) / d[:, k].astype(np.float64)
        cell_t = np.abs(float(ext[k]) / 32768.0 / d[:, k].astype(np.float64))
        assert np.all(np.abs(t32 - t64) <= 0.05 * cell_t), float(np.max(np.abs(t32 - t64) / cell_t))


def test_entry_plane_selector_equals_min_of_both_planes():
    rng = np.random.default_rng(3)
    n = 5000
    qlo = rng.integers(0, 32000, n).astype(np.uint32)
    qhi = (qlo + rng.integers(0, 700, n)).astype(np.uint32)
    words = (0x8000 | qlo) | ((0x8000 | qhi) << 16)
    inv = rng.normal(size=n).astype(F) * F(100.0)
    oid = rng.normal(size=n).astype(F) * F(50.0)
    t_lo = (decode(words, SEL_LOW).astype(np.float64) * inv + oid).astype(F)
    t_hi = (decode(words, SEL_HIGH).astype(np.float64) * inv + oid).astype(F)
    near_sel = np.where(inv < 0, t_hi, t_lo)
    far_sel = np.where(inv < 0, t_lo, t_hi)
    assert np.array_equal(near_sel, np.minimum(t_lo, t_hi)) and np.array_equal(far_sel, np.maximum(t_lo, t_hi))
