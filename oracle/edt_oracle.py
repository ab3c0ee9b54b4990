"""CPU oracle for the exact Euclidean distance transform / signed distance map.

TEST INFRASTRUCTURE — never imported by the product package.

The reference calls ``scipy.ndimage.distance_transform_edt`` (a third-party dependency that is
not vendored under /root/reference and is version-pinned nowhere; scipy 1.18.1 is what this
image carries) from ``src/train_with_boundary_loss.py:197-198``.  Its published semantics:
for every non-zero input pixel, the Euclidean distance to the nearest zero pixel, float64;
zero pixels map to 0.  The result is mathematically unique, so any exact algorithm is an
oracle.  This file restates it as the classic two-pass separable scheme in *integers*
(squared distances), followed by one float64 sqrt — and ``tests/test_oracle_golden.py`` pins
it against scipy itself (run through the reference's own ``signed_distance_map_np``) on the
golden masks.
"""
from __future__ import annotations

import numpy as np

_INF = np.int64(1) << 20          # larger than any in-image distance, (_INF)^2 fits int64


def _column_distance(zero: np.ndarray) -> np.ndarray:
    """g[y,x] = distance along column x from row y to the nearest row whose pixel is 'zero'
    (a feature); _INF if the column has none."""
    H, W = zero.shape
    g = np.full((H, W), _INF, dtype=np.int64)
    run = np.full(W, _INF, dtype=np.int64)
    for y in range(H):                              # downward sweep
        run = np.where(zero[y], 0, np.minimum(run + 1, _INF))
        g[y] = run
    run = np.full(W, _INF, dtype=np.int64)
    for y in range(H - 1, -1, -1):                  # upward sweep
        run = np.where(zero[y], 0, np.minimum(run + 1, _INF))
        g[y] = np.minimum(g[y], run)
    return g


def edt_squared(nonzero: np.ndarray) -> np.ndarray:
    """Exact squared EDT (int64): for pixels where ``nonzero`` is True, the squared distance to
    the nearest False pixel; 0 where ``nonzero`` is False.  Requires at least one False pixel."""
    nonzero = np.asarray(nonzero, dtype=bool)
    H, W = nonzero.shape
    g = _column_distance(~nonzero)
    g2 = g * g
    xs = np.arange(W, dtype=np.int64)
    dx2 = (xs[:, None] - xs[None, :]) ** 2          # [x, x']
    out = np.empty((H, W), dtype=np.int64)
    step = max(1, (1 << 24) // max(1, W * W))
    for y0 in range(0, H, step):
        blk = g2[y0:y0 + step]                      # [h, x']
        out[y0:y0 + step] = (dx2[None, :, :] + blk[:, None, :]).min(axis=2)
    return out


def edt(nonzero: np.ndarray) -> np.ndarray:
    """float64 EDT with scipy.ndimage.distance_transform_edt semantics (2-D, unit sampling)."""
    return np.sqrt(edt_squared(nonzero).astype(np.float64))


def sdf_of_mask(mask: np.ndarray) -> np.ndarray:
    """src/train_with_boundary_loss.py:191-202 — SDF, negative inside, positive outside, float32;
    all-foreground / all-background masks give all zeros."""
    mask = np.asarray(mask).astype(bool)
    if mask.any() and (~mask).any():
        return (edt(~mask) - edt(mask)).astype(np.float32)
    return np.zeros(mask.shape, dtype=np.float32)


def edt_truncated_int(mask: np.ndarray) -> np.ndarray:
    """src/training/losses/abl.py:17-24 behaviour ("next" row N1): distance map assigned into an
    int32 array, i.e. truncated toward zero."""
    return edt(mask).astype(np.int32)
