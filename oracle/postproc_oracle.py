"""CPU oracle for the pseudo-label post-processing ("next" row N3 of SURVEY.md §8f).

TEST INFRASTRUCTURE — never imported by the product package.

Restates
  * ``ensemble_forward``   src/data_preprocessing/create_pseudo_labels_gpu.py:201-215  (sum_m w_m * sigmoid(logits_m))
  * threshold + QC scores  :294-301 with ``entropy_map`` :128-130  (fg area, median confidence, mean entropy)
  * ``should_accept``      :141-147
  * ``clean_mask``         src/data_preprocessing/clean_masks.py:12-32  (flood-fill hole filling from pixel (0,0), then
                           the largest 8-connected component)
  * ``clean_mask_largest_component``  src/data_preprocessing/remove_blops.py:14-33
in numpy/scipy.  Third-party arithmetic: the reference calls OpenCV (``cv2.floodFill``, 4-connected by default;
``cv2.connectedComponentsWithStats(connectivity=8)``), unpinned in the repository; opencv 4.13.0 is what this image
carries.  Connected components are mathematically unique; the one implementation-defined choice is WHICH component
wins when several share the largest area: ``1 + argmax(stats[1:, AREA])`` takes the lowest label, and OpenCV's
block-based labelling numbers components by the first 2x2 block (block-row major) that touches them — reproduced
by ``_cv_label_order_key`` and pinned against OpenCV itself in tests/golden/postproc.npz.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

_S8 = np.ones((3, 3), dtype=bool)
_S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def ensemble_probs(logits_list, weights) -> np.ndarray:
    """create_pseudo_labels_gpu.py:167-171,201-215 — weights normalised to sum 1 (float32), probs accumulated in fp32
    (the reference does it in fp16 under autocast; the build keeps fp32, as for every other map)."""
    w = np.asarray(weights, dtype=np.float32)
    w = w / w.sum()
    out = None
    for z, wi in zip(logits_list, w):
        p = sigmoid_f32(np.asarray(z)) * np.float32(wi)
        out = p if out is None else out + p
    return out.astype(np.float32)


def entropy_map(p: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """create_pseudo_labels_gpu.py:128-130 in float32 (numpy keeps float32 with Python-float bounds)."""
    p = np.clip(p.astype(np.float32), np.float32(eps), np.float32(1 - eps))
    return -(p * np.log(p) + (np.float32(1) - p) * np.log(np.float32(1) - p))


def qc_scores(probs: np.ndarray, threshold: float = 0.5):
    """:294-299 for one [H,W] float32 probability map -> (pred01 uint8, fg_area, fg_conf, mean_entropy)."""
    probs = np.asarray(probs, dtype=np.float32)
    pred01 = (probs >= threshold).astype(np.uint8)
    fg_area = float(pred01.mean())
    fg_conf = float(np.median(np.abs(probs - np.float32(0.5)) * np.float32(2.0)))
    mean_ent = float(entropy_map(probs).mean())
    return pred01, fg_area, fg_conf, mean_ent


def should_accept(fg_area, fg_conf, mean_entropy, tta_iou=1.0, edge_hit=1.0, *, min_fg_area=0.005, max_fg_area=0.60,
                  min_fg_conf=0.65, max_mean_ent=0.35, enable_tta_iou=False, min_tta_iou=0.75, min_edge_hit=0.10) -> bool:
    """:141-147 with the thresholds of :58-64."""
    if fg_area < min_fg_area or fg_area > max_fg_area:
        return False
    if fg_conf < min_fg_conf:
        return False
    if mean_entropy > max_mean_ent:
        return False
    if enable_tta_iou and tta_iou < min_tta_iou:
        return False
    if edge_hit < min_edge_hit:
        return False
    return True


def _cv_label_order_key(labels: np.ndarray, n: int) -> np.ndarray:
    """For components 1..n: position, in block-row-major order of 2x2 blocks, of the first block touching them."""
    H, W = labels.shape
    yy, xx = np.mgrid[0:H, 0:W]
    key = (yy >> 1) * ((W + 1) >> 1) + (xx >> 1)
    out = np.full(n + 1, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(out, labels.ravel(), key.ravel())
    return out[1:]


def largest_component(fg: np.ndarray) -> np.ndarray:
    """bool [H,W] -> bool [H,W]: the largest 8-connected component (ties: OpenCV's lowest label); unchanged when
    there is no foreground (clean_masks.py:26-27, remove_blops.py:26-27)."""
    fg = np.asarray(fg, dtype=bool)
    lab, n = ndimage.label(fg, structure=_S8)
    if n == 0:
        return fg.copy()
    area = np.bincount(lab.ravel(), minlength=n + 1)[1:]
    order = _cv_label_order_key(lab, n)
    best = min(range(n), key=lambda i: (-int(area[i]), int(order[i])))
    return lab == best + 1


def fill_holes_from_corner(fg: np.ndarray) -> np.ndarray:
    """clean_masks.py:16-22: flood-fill the background from pixel (0,0) (4-connected); whatever background the fill
    did not reach becomes foreground.  If (0,0) is itself foreground the fill changes nothing and the inverse of the
    'filled' image is the whole background: the result is all-foreground."""
    fg = np.asarray(fg, dtype=bool)
    if fg[0, 0]:
        return np.ones_like(fg)
    lab, _ = ndimage.label(~fg, structure=_S4)
    reached = lab == lab[0, 0]
    return fg | (~fg & ~reached)


def clean_mask(mask_u8: np.ndarray) -> np.ndarray:
    """clean_masks.py:12-32 — uint8 {0,255} out."""
    binary = np.asarray(mask_u8) > 127
    clean = fill_holes_from_corner(binary)
    return (largest_component(clean).astype(np.uint8)) * np.uint8(255)


def clean_mask_largest_component(mask_u8: np.ndarray) -> np.ndarray:
    """remove_blops.py:14-33 — {0,255} out, except that a mask without foreground is returned as {0,1} zeros."""
    binary = np.asarray(mask_u8) > 0
    return (largest_component(binary).astype(np.uint8)) * np.uint8(255)
