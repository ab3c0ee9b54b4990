"""CPU oracle for the Active Boundary Loss ("next" row N1 of SURVEY.md §8f).

TEST INFRASTRUCTURE — never imported by the product package.

Restates, for the binary (one logit channel) case the reference trains with,
  * ``ABL.forward`` and helpers            src/training/losses/abl.py:14-30, 66-212
  * ``LabelSmoothSoftmaxCEV1`` (reduction 'none')   src/training/losses/label_smooth.py:14-57
  * ``BCEDiceABL.components``              src/training/train_BCEDice_ABL.py:264-298
as a dense per-pixel formulation (no ``nonzero`` gather lists): every quantity is a full [B,H,W] map and
the selection of boundary pixels is a mask.  Pinned against the reference's own class by
``tests/test_oracle_golden.py`` on ``tests/golden/abl.npz`` (``oracle/make_golden.py``).

Behaviours of the reference that are reproduced deliberately (they are what "identical results" means):
  1. the two-channel "logits" it feeds to softmax / KL are the PROBABILITIES (1-p, p) (abl.py:186-189);
  2. the GT distance map is the EDT assigned into an int32 array, i.e. truncated toward zero (abl.py:18-23),
     minus one, clamped at zero (abl.py:168-169);
  3. ``get_dist_maps`` concatenates the per-image [2,H,W] maps along dim 0 (abl.py:166-167), so the map the
     loss looks up for batch entry n is channel (n % 2) of image (n // 2) — not image n's own map.
     ``per_image_maps=True`` switches to the evidently intended behaviour (channel 0 of image n);
  4. scipy's EDT of an input without any zero (a GT with no boundary at all) measures distances to the
     out-of-image position (row -1, column 0) — scipy 1.18.1, reproduced by ``_edt_sq_vs``;
  5. the adaptive KL threshold (abl.py:78-83) counts pixels over the whole batch against H*W/100;
  6. label smoothing puts 1-s on the target direction and s/8 on all 8 (label_smooth.py:43-45), so the
     smoothed distribution sums to 1 - s + s = 0.975 + ... (0.8 + 8*0.025 - 0.025 = 0.975 for s = 0.2).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .edt_oracle import edt_squared

Tensor = torch.Tensor

# (d_row, d_col) of the nine candidate directions in the reference's order (abl.py:126-133); index 8 = stay
DIRECTIONS = ((1, 0), (-1, 0), (0, -1), (0, 1), (-1, 1), (1, 1), (-1, -1), (1, -1), (0, 0))
MAX_DIS = 1e5            # padding value of the distance map outside the image (abl.py:116,120)


def eps_ladder(n: int = 80) -> np.ndarray:
    """The thresholds the ``while True`` loop of abl.py:78-83 can visit: eps_0 = 1e-5, eps_{k+1} = eps_k * 1.2 in
    Python floats (float64); the comparison ``kl > eps`` is made in float32."""
    out, e = [], 1e-5
    for _ in range(n):
        out.append(e)
        e *= 1.2
    return np.asarray(out, dtype=np.float64).astype(np.float32)


def gt_boundary(labels: np.ndarray, ignore_label: int = 255) -> np.ndarray:
    """abl.py:94-107 — pixel differs from the one below or the one to the right, or carries the ignore label."""
    lab = np.asarray(labels)
    gb = np.zeros(lab.shape, dtype=bool)
    gb[:, :-1, :] |= lab[:, 1:, :] != lab[:, :-1, :]
    gb[:, :, :-1] |= lab[:, :, 1:] != lab[:, :, :-1]
    gb |= lab == ignore_label
    return gb


def _edt_sq_vs(nonzero: np.ndarray) -> np.ndarray:
    """Squared EDT of ``nonzero`` with scipy's behaviour when there is no zero pixel at all."""
    if (~nonzero).any():
        return edt_squared(nonzero)
    H, W = nonzero.shape
    yy, xx = np.mgrid[0:H, 0:W]
    return ((yy + 1) ** 2 + xx ** 2).astype(np.int64)


def _isqrt(a: np.ndarray) -> np.ndarray:
    r = np.floor(np.sqrt(a.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > a, r - 1, r)
    return np.where((r + 1) * (r + 1) <= a, r + 1, r)


def one_hot_dist_channel(gb: np.ndarray, channel: int) -> np.ndarray:
    """abl.py:16-24 + 168-169 for one image: ``max(0, -one_hot2dist(class2one_hot(gb))[channel])`` as float32.
    channel 0: positive class = non-boundary pixels; channel 1: positive class = boundary pixels."""
    pos = ~gb if channel == 0 else gb
    out = np.zeros(gb.shape, dtype=np.float32)
    if not pos.any():
        return out
    d = _isqrt(_edt_sq_vs(pos)) - 1            # trunc(-(d - 1)) = -(floor(d) - 1) for d >= 1
    out[pos] = np.maximum(d[pos], 0).astype(np.float32)
    return out


def dist_maps(gb: np.ndarray, per_image_maps: bool = False) -> np.ndarray:
    """The [B,H,W] map the loss indexes with the batch index (see behaviour 3 in the module docstring)."""
    B = gb.shape[0]
    if per_image_maps:
        return np.stack([one_hot_dist_channel(gb[n], 0) for n in range(B)])
    return np.stack([one_hot_dist_channel(gb[n // 2], n % 2) for n in range(B)])


def _kl(center: Tensor, other: Tensor) -> Tensor:
    """abl.py:14-15 ``kl_div(a=center, b=other)`` summed over the channel dim (dim 1)."""
    return (F.softmax(other, dim=1) * (F.log_softmax(other, dim=1) - F.log_softmax(center, dim=1))).sum(1)


def pred_boundary(probs2: Tensor, max_n_ratio: float = 0.01) -> Tuple[Tensor, int, Tensor]:
    """abl.py:66-92.  Returns (dilated boundary mask [B,H,W] bool, ladder index used, kl map [B,H,W])."""
    B, _, H, W = probs2.shape
    kl = torch.zeros(B, H, W, dtype=probs2.dtype)
    kl[:, :-1, :] += _kl(probs2[:, :, 1:, :], probs2[:, :, :-1, :])
    kl_lr = torch.zeros(B, H, W, dtype=probs2.dtype)
    kl_lr[:, :, :-1] = _kl(probs2[:, :, :, 1:], probs2[:, :, :, :-1])
    kl = kl_lr + kl                                        # reference adds lr + ud in this order
    max_n = float(np.float32(H * W * max_n_ratio))        # torch compares a float32 tensor with a Python scalar in float32
    ladder = eps_ladder()
    k = 0
    while float((kl > float(ladder[k])).sum()) > max_n:
        k += 1
    core = (kl > float(ladder[k])).float()[:, None]
    dil = F.max_pool2d(core, 3, stride=1, padding=1)[:, 0] > 0
    return dil, k, kl


def abl_loss(logits: Tensor, target: Tensor, label_smoothing: float = 0.2, max_n_ratio: float = 0.01,
             max_clip_dist: float = 20.0, ignore_label: int = 255, per_image_maps: bool = False,
             return_parts: bool = False):
    """ABL.forward for [B,1,H,W] logits.  Returns the scalar loss tensor (differentiable w.r.t. ``logits``) or
    ``None`` when the predicted boundary is empty (abl.py:197-198)."""
    assert logits.dim() == 4 and logits.shape[1] == 1
    tgt = target[:, 0] if target.dim() == 4 else target
    labels = tgt.long().numpy()
    B, _, H, W = logits.shape
    p = torch.sigmoid(logits)
    probs2 = torch.cat([1.0 - p, p], dim=1)                                   # abl.py:186-189
    gb = gt_boundary(labels, ignore_label)
    dmap = torch.from_numpy(dist_maps(gb, per_image_maps))                    # [B,H,W] float32
    pb, k, kl_map = pred_boundary(probs2.detach(), max_n_ratio)
    if int(pb.sum()) < 1:
        return (None, dict(pred_boundary=pb, k=k)) if return_parts else None

    dpad = F.pad(dmap, (1, 1, 1, 1), value=MAX_DIS)
    ppad = F.pad(probs2, (1, 1, 1, 1), mode="replicate")
    cand = torch.stack([dpad[:, 1 + dr:1 + dr + H, 1 + dc:1 + dc + W] for dr, dc in DIRECTIONS], 0)   # [9,B,H,W]
    direction = torch.argmin(cand, dim=0)                                     # first minimum wins
    keep = pb & (direction != 8)
    kls = torch.stack([_kl(probs2, ppad[:, :, 1 + dr:1 + dr + H, 1 + dc:1 + dc + W].detach())
                       for dr, dc in DIRECTIONS[:8]], 1)                      # [B,8,H,W]
    logs = F.log_softmax(kls, dim=1)
    lb = torch.full_like(kls, label_smoothing / 8.0)
    lb.scatter_(1, direction.clamp(max=7)[:, None], 1.0 - label_smoothing)
    ce = -(logs * lb).sum(1)                                                  # [B,H,W]
    weight = torch.clamp(dmap, max=max_clip_dist) / max_clip_dist
    n_keep = int(keep.sum())
    loss = (ce * weight)[keep].sum() / n_keep if n_keep else (ce * weight)[keep].mean()
    if return_parts:
        return loss, dict(pred_boundary=pb, k=k, direction=direction, keep=keep, dmap=dmap, gt_boundary=gb,
                          kl_map=kl_map, n_keep=n_keep)
    return loss


def bce_dice_abl(logits: Tensor, target: Tensor, bce_weight: float = 0.5, smooth: float = 1.0,
                 abl_weight: float = 0.1, **abl_kwargs) -> Tensor:
    """BCEDiceABL.forward — src/training/train_BCEDice_ABL.py:264-298: region + abl_weight * boundary, or the region
    term alone when ABL returns None."""
    from .unet_oracle import bce_dice_loss
    region = bce_dice_loss(logits, target, bce_weight, smooth)
    boundary = abl_loss(logits, target, **abl_kwargs)
    return region if boundary is None else region + abl_weight * boundary
