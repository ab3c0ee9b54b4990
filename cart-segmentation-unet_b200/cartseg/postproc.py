"""Pseudo-label post-processing on the device.

  ensemble_forward                 src/data_preprocessing/create_pseudo_labels_gpu.py:201-215
  pseudo_label_qc, should_accept   :141-147, 294-300  (mask, foreground area, median confidence, mean entropy)
  clean_mask                       src/data_preprocessing/clean_masks.py:12-32
  clean_mask_largest_component     src/data_preprocessing/remove_blops.py:14-33

The reference copies 4 B/px of probabilities to the host and computes these with numpy / OpenCV per image; here the
probabilities never leave the GPU: 1 B/px of mask and three numbers per image come back.  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import CartsegError, check, ptr


def _f32c(t: Tensor) -> Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


@torch.library.custom_op("cartseg::ensemble_accumulate", mutates_args=("probs",), device_types="cuda")
def ensemble_accumulate(logits: Tensor, weight: float, first: bool, probs: Tensor) -> None:
    if not (logits.is_cuda and probs.is_cuda):
        raise CartsegError("cartseg::ensemble_accumulate takes CUDA tensors only (no CPU fallback)")
    if logits.dtype != torch.float32 or probs.dtype != torch.float32 or logits.numel() != probs.numel():
        raise CartsegError("ensemble_accumulate: float32 logits / probs of equal size required")
    if not (logits.is_contiguous() and probs.is_contiguous()):
        raise CartsegError("ensemble_accumulate: contiguous tensors required")
    with torch.cuda.device(logits.device):
        check(_lib.lib().cs_ensemble_accumulate(ptr(logits), float(weight), logits.numel(), int(first), ptr(probs),
                                                _lib.current_stream()), "cs_ensemble_accumulate")


@torch.library.custom_op("cartseg::pseudo_qc", mutates_args=(), device_types="cuda")
def pseudo_qc(probs: Tensor, threshold: float, mask_value: int) -> Tuple[Tensor, Tensor]:
    """probs [B,H,W] float32 -> (mask uint8 [B,H,W] in {0, mask_value}, stats float64 [B,4] =
    {foreground pixels, median(|p-0.5|*2), mean entropy, pixels})."""
    if not probs.is_cuda:
        raise CartsegError("cartseg::pseudo_qc takes CUDA tensors only (no CPU fallback)")
    if probs.dim() != 3 or probs.dtype != torch.float32 or not probs.is_contiguous():
        raise CartsegError("pseudo_qc: probs must be contiguous float32 [B,H,W]")
    B, H, W = probs.shape
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=probs.device)
    stats = torch.empty((B, 4), dtype=torch.float64, device=probs.device)
    with torch.cuda.device(probs.device):
        check(_lib.lib().cs_pseudo_qc(ptr(probs), B, H * W, float(threshold), int(mask_value), ptr(mask), ptr(stats),
                                      _lib.current_stream()), "cs_pseudo_qc")
    return mask, stats


@pseudo_qc.register_fake
def _(probs, threshold, mask_value):
    return probs.new_empty(probs.shape, dtype=torch.uint8), probs.new_empty((probs.shape[0], 4), dtype=torch.float64)


@torch.library.custom_op("cartseg::mask_cleanup", mutates_args=(), device_types="cuda")
def mask_cleanup(mask: Tensor, bin_threshold: int, fill_holes: bool, keep_largest: bool) -> Tensor:
    """mask uint8 [B,H,W] (or [H,W]) -> uint8 {0,255} of the same shape."""
    if not mask.is_cuda:
        raise CartsegError("cartseg::mask_cleanup takes CUDA tensors only (no CPU fallback)")
    if mask.dtype != torch.uint8 or not mask.is_contiguous() or mask.dim() not in (2, 3):
        raise CartsegError("mask_cleanup: mask must be contiguous uint8 [B,H,W] or [H,W]")
    B = mask.shape[0] if mask.dim() == 3 else 1
    H, W = mask.shape[-2], mask.shape[-1]
    out = torch.empty_like(mask)
    L = _lib.lib()
    scratch = torch.empty(int(L.cs_mask_cleanup_scratch_bytes(B, H, W)), dtype=torch.uint8, device=mask.device)
    with torch.cuda.device(mask.device):
        check(L.cs_mask_cleanup(ptr(mask), B, H, W, int(bin_threshold), int(fill_holes), int(keep_largest), ptr(out),
                                ptr(scratch), _lib.current_stream()), "cs_mask_cleanup")
    return out


@mask_cleanup.register_fake
def _(mask, bin_threshold, fill_holes, keep_largest):
    return torch.empty_like(mask)


# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def ensemble_forward(models: Sequence, weights: Sequence[float], tens: Tensor) -> Tensor:
    """create_pseudo_labels_gpu.py:201-215: ``sum_m w_m * sigmoid(model_m(tens))[:, 0]`` -> [B,H,W] float32.
    ``weights`` are used as given (load_ensemble normalises them, :167-171 — see :func:`normalize_weights`)."""
    if len(models) != len(weights) or not models:
        raise CartsegError("ensemble_forward: one weight per model required")
    probs = None
    for i, (m, w) in enumerate(zip(models, weights)):
        logits = m(tens)
        if not logits.is_cuda:
            raise CartsegError("ensemble_forward takes CUDA tensors only (no CPU fallback)")
        if logits.dim() != 4 or logits.shape[1] != 1:
            raise CartsegError("ensemble_forward: models must return [B,1,H,W] logits")
        lg = _f32c(logits)
        if probs is None:
            probs = torch.empty((lg.shape[0], lg.shape[2], lg.shape[3]), dtype=torch.float32, device=lg.device)
        torch.ops.cartseg.ensemble_accumulate(lg, float(w), i == 0, probs)
    return probs


def normalize_weights(weights: Sequence[float]) -> List[float]:
    """load_ensemble, create_pseudo_labels_gpu.py:167-171: float32 weights divided by their float32 sum."""
    import numpy as np
    w = np.array(weights, dtype=np.float32)
    return (w / w.sum()).tolist()


@torch.no_grad()
def pseudo_label_qc(probs: Tensor, threshold: float = 0.5, mask_value: int = 1):
    """create_pseudo_labels_gpu.py:294-300 for a whole batch.  Returns ``(pred01, fg_area, fg_conf, mean_entropy)``:
    the uint8 mask [B,H,W] and three float64 [B] tensors, all on the device (one D2H of 3 numbers per image instead
    of 4 B/px)."""
    if not probs.is_cuda:
        raise CartsegError("pseudo_label_qc takes CUDA tensors only (no CPU fallback)")
    if probs.dim() == 4 and probs.shape[1] == 1:
        probs = probs[:, 0]
    mask, stats = torch.ops.cartseg.pseudo_qc(_f32c(probs), float(threshold), int(mask_value))
    return mask, stats[:, 0] / stats[:, 3], stats[:, 1], stats[:, 2]


def should_accept(fg_area: float, fg_conf: float, mean_entropy: float, tta_iou: float = 1.0, edge_hit: float = 1.0, *,
                  min_fg_area: float = 0.005, max_fg_area: float = 0.60, min_fg_conf: float = 0.65,
                  max_mean_ent: float = 0.35, enable_tta_iou: bool = False, min_tta_iou: float = 0.75,
                  min_edge_hit: float = 0.10) -> bool:
    """create_pseudo_labels_gpu.py:141-147 with the thresholds of :58-64 as defaults (host logic, unchanged)."""
    if fg_area < min_fg_area or fg_area > max_fg_area:
        return False
    if fg_conf < min_fg_conf:
        return False
    if mean_entropy > max_mean_ent:
        return False
    if enable_tta_iou and (tta_iou < min_tta_iou):
        return False
    if edge_hit < min_edge_hit:
        return False
    return True


def _u8c(mask: Tensor) -> Tensor:
    if not mask.is_cuda:
        raise CartsegError("mask clean-up takes CUDA tensors only (no CPU fallback)")
    return mask if (mask.dtype == torch.uint8 and mask.is_contiguous()) else mask.to(torch.uint8).contiguous()


@torch.no_grad()
def clean_mask(mask: Tensor) -> Tensor:
    """clean_masks.py:12-32: binarise at > 127, fill the holes the flood fill from (0,0) cannot reach, keep the largest
    8-connected component.  uint8 {0,255}, [B,H,W] or [H,W]."""
    return torch.ops.cartseg.mask_cleanup(_u8c(mask), 127, True, True)


@torch.no_grad()
def clean_mask_largest_component(mask: Tensor) -> Tensor:
    """remove_blops.py:14-33: binarise at > 0, keep the largest 8-connected component.  uint8 {0,255} (a mask without
    any foreground comes back as zeros; the reference returns its {0,1} zeros there)."""
    return torch.ops.cartseg.mask_cleanup(_u8c(mask), 0, False, True)
