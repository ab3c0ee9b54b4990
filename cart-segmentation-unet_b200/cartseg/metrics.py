"""Metrics and thresholding with the reference's function names.

  dice_metric, iou_metric, find_best_threshold   train_bce_dice.py:201-232
  precision_recall_f1                             src/train_with_focalDice.py:266-284
  hard_dice_metric, hard_iou_metric, dice_iou_at_t, sweep_thresholds
                                                  src/finetune_pseudo.py:192-226, src/finetune_for_224.py:223-248
  pseudo_label_mask                               src/data_preprocessing/create_pseudo_labels_gpu.py:212,294

One pass over (logits, targets) yields, per sample and per threshold, the sums every one of these
formulas needs (``cartseg::threshold_stats``); the reference's threshold sweeps re-run the model once
per threshold, here one forward serves all thresholds.  Thresholding happens in
logit space with the exact float32 bound (ops.logit_bound), so masks equal ``sigmoid(x) > t`` bit for bit.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
from torch import Tensor

from . import ops
from ._lib import CartsegError


def _prep(logits: Tensor, targets: Tensor) -> Tuple[Tensor, Tensor, int]:
    if not logits.is_cuda:
        raise CartsegError("cartseg metrics take CUDA tensors only (no CPU fallback)")
    if logits.shape != targets.shape or logits.dim() != 4:
        raise CartsegError("expected logits and targets of the same [B,C,H,W] shape")
    lg = logits.detach()
    lg = lg if (lg.dtype == torch.float32 and lg.is_contiguous()) else lg.float().contiguous()
    tg = targets.detach()
    tg = tg if (tg.dtype == torch.float32 and tg.is_contiguous()) else tg.float().contiguous()
    return lg, tg, lg.shape[0] * lg.shape[1]


@torch.no_grad()
def threshold_sums(logits: Tensor, targets: Tensor, thresholds: Sequence[float], ge: bool = False):
    """Per (b,c) row: pred_sum[k], inter[k] for every threshold, plus target_sum and the soft sums.

    Returns (pred_sum [R,K], inter [R,K], target_sum [R], p_sum [R], pt_sum [R]) as float64 CUDA tensors."""
    lg, tg, rows = _prep(logits, targets)
    xs = torch.tensor([ops.logit_bound(float(t), ge) for t in thresholds], dtype=torch.float32, device=lg.device)
    counts, soft = torch.ops.cartseg.threshold_stats(lg, tg, rows, xs)
    return counts[:, :, 0], counts[:, :, 1], soft[:, 1], soft[:, 0], soft[:, 2]


@torch.no_grad()
def dice_metric(logits: Tensor, targets: Tensor, smooth: float = 1.0, eps: float = 1e-7) -> float:
    """Soft Dice, train_bce_dice.py:201-206."""
    _, _, t_sum, p_sum, pt_sum = threshold_sums(logits, targets, [0.5])
    return ((2 * pt_sum + smooth) / (p_sum + t_sum + smooth + eps)).mean().item()


@torch.no_grad()
def iou_metric(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> float:
    """Hard IoU, train_bce_dice.py:208-212."""
    ps, inter, ts, _, _ = threshold_sums(logits, targets, [t])
    ps, inter = ps[:, 0], inter[:, 0]
    return ((inter + eps) / (ps + ts - inter + eps)).mean().item()


def _per_sample(v: Tensor, B: int) -> Tensor:
    return v.reshape(B, -1, *v.shape[1:]).sum(1)          # dims (1,2,3): fold the channel rows


@torch.no_grad()
def dice_iou_at_t(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> Tuple[float, float]:
    """src/finetune_for_224.py:223-233 (sums over dims (1,2,3))."""
    B = logits.shape[0]
    ps, inter, ts, _, _ = threshold_sums(logits, targets, [t])
    ps, inter, ts = _per_sample(ps[:, 0], B), _per_sample(inter[:, 0], B), _per_sample(ts, B)
    den = ps + ts
    dice = (2 * inter + eps) / (den + eps)
    iou = (inter + eps) / (den - inter + eps)
    return float(dice.mean().item()), float(iou.mean().item())


def hard_dice_metric(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> float:
    """src/finetune_pseudo.py:192-199."""
    return dice_iou_at_t(logits, targets, t, eps)[0]


def hard_iou_metric(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> float:
    """src/finetune_pseudo.py:201-208."""
    return dice_iou_at_t(logits, targets, t, eps)[1]


@torch.no_grad()
def precision_recall_f1(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7):
    """src/train_with_focalDice.py:266-284."""
    ps, tp, ts, _, _ = threshold_sums(logits, targets, [t])
    ps, tp = ps[:, 0], tp[:, 0]
    fp, fn = ps - tp, ts - tp
    prec = tp / (tp + fp + eps)
    rec = tp / (tp + fn + eps)
    f1 = 2 * prec * rec / (prec + rec + eps)
    return tuple(torch.nan_to_num(v).mean().item() for v in (prec, rec, f1))


@torch.no_grad()
def sweep_thresholds(logits: Tensor, targets: Tensor, thresholds: Iterable[float], smooth: float = 1.0) -> Tensor:
    """Batch-mean hard Dice ``(2I + smooth) / (P + T + smooth)`` for every threshold from ONE pass
    (the inner expression of find_best_threshold, train_bce_dice.py:223-227).  Returns a float64 [K] tensor."""
    ths = [float(t) for t in thresholds]
    ps, inter, ts, _, _ = threshold_sums(logits, targets, ths)
    return ((2 * inter + smooth) / (ps + ts[:, None] + smooth)).mean(0)


@torch.no_grad()
def find_best_threshold(model, val_loader, device, thresholds=None):
    """train_bce_dice.py:214-232 with one forward per batch instead of one per (batch, threshold)."""
    import numpy as np
    if thresholds is None:
        thresholds = np.linspace(0.2, 0.8, 13)
    ths = [float(t) for t in thresholds]
    model.eval()
    per_batch: List[Tensor] = []
    for data, target in val_loader:
        data, target = data.to(device), target.to(device)
        per_batch.append(sweep_thresholds(model(data), target, ths, smooth=1.0))
    md = torch.stack(per_batch).mean(0).tolist()
    best_t, best_d = 0.5, -1.0
    for t, d in zip(thresholds, md):
        if d > best_d:
            best_d, best_t = d, t
    return best_t, best_d


@torch.no_grad()
def pseudo_label_mask(logits: Tensor, threshold: float = 0.5) -> Tensor:
    """uint8 [B,H,W] mask = sigmoid(logits)[:,0] >= threshold (create_pseudo_labels_gpu.py:212,294),
    produced on the device at 1 B/px instead of shipping float32 probabilities to the host."""
    if logits.dim() != 4 or logits.shape[1] != 1:
        raise CartsegError("expected [B,1,H,W] logits")
    lg = logits.detach()
    lg = lg if (lg.dtype == torch.float32 and lg.is_contiguous()) else lg.float().contiguous()
    return torch.ops.cartseg.threshold_mask(lg, ops.logit_bound(float(threshold), ge=True))[:, 0]
