"""Image-parity statistics: per-pixel z-scores between two Monte-Carlo renders, tile z-scores, RMSE / PSNR.

Pure numpy helpers; they take sums / sums of squares and know nothing about where those came from.
"""
from __future__ import annotations

import numpy as np


def z_scores(sum_a, sumsq_a, n_a, sum_b, sumsq_b, n_b):
    """Per pixel and channel z = (m_A - m_B) / sqrt(s2_A/n_A + s2_B/n_B).  Entries without usable variance are masked out:
    exactly zero, or below (1 % of the mean)^2 — for (nearly) deterministic pixels such as a constant background the
    float32 sums of squares only hold accumulation noise, and the ratio of two rounding errors is not a z-score."""
    sum_a, sumsq_a, sum_b, sumsq_b = (np.asarray(x, np.float64) for x in (sum_a, sumsq_a, sum_b, sumsq_b))
    m_a, m_b = sum_a / n_a, sum_b / n_b
    v_a = np.maximum(sumsq_a / n_a - m_a * m_a, 0.0) * n_a / max(n_a - 1, 1)
    v_b = np.maximum(sumsq_b / n_b - m_b * m_b, 0.0) * n_b / max(n_b - 1, 1)
    se2 = v_a / n_a + v_b / n_b
    valid = (se2 > 0) & ((v_a + v_b) > 1e-4 * np.maximum(np.abs(m_a), np.abs(m_b)) ** 2)
    z = np.zeros_like(m_a)
    z[valid] = (m_a[valid] - m_b[valid]) / np.sqrt(se2[valid])
    return z, valid


def tile_z_scores(sum_a, sumsq_a, n_a, sum_b, sumsq_b, n_b, tile: int = 50):
    """z-scores of tile-mean differences, one per (tile, channel).  Channels are NOT pooled: the three channels of a
    sample are strongly correlated (same path), so a pooled variance would be underestimated by up to 3x."""
    sum_a, sumsq_a, sum_b, sumsq_b = (np.asarray(x, np.float64) for x in (sum_a, sumsq_a, sum_b, sumsq_b))
    h, w, _ = sum_a.shape
    m_a, m_b = sum_a / n_a, sum_b / n_b
    v_a = np.maximum(sumsq_a / n_a - m_a * m_a, 0.0) / max(n_a - 1, 1)  # variance of the pixel mean
    v_b = np.maximum(sumsq_b / n_b - m_b * m_b, 0.0) / max(n_b - 1, 1)
    out = []
    for y in range(0, h - tile + 1, tile):
        for x in range(0, w - tile + 1, tile):
            sl = (slice(y, y + tile), slice(x, x + tile))
            diff = (m_a[sl] - m_b[sl]).sum(axis=(0, 1))
            var = (v_a[sl] + v_b[sl]).sum(axis=(0, 1))
            out.append(np.where(var > 0, diff / np.sqrt(np.maximum(var, 1e-300)), 0.0))
    return np.array(out)


def summary(z, valid):
    zz = z[valid]
    n = zz.size
    return {"n": int(n), "mean_z": float(zz.mean()) if n else 0.0, "std_z": float(zz.std()) if n else 0.0,
            "frac_gt3": float((np.abs(zz) > 3).mean()) if n else 0.0, "frac_gt4": float((np.abs(zz) > 4).mean()) if n else 0.0}


def rmse_psnr(img, ref, peak: float = 1.0):
    """RMSE and PSNR of `img` against `ref` after clamping both to [0, peak] (display range)."""
    a = np.clip(np.asarray(img, np.float64), 0, peak)
    b = np.clip(np.asarray(ref, np.float64), 0, peak)
    mse = float(((a - b) ** 2).mean())
    rmse = mse ** 0.5
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
    return rmse, psnr
