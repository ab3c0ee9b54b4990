"""CPU oracle for the U-Net hot path (TEST INFRASTRUCTURE — never imported by the product).

A functional, fp32/fp64 PyTorch-CPU restatement of the reference's model, losses and
metrics.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module; the product package (``cartseg``) fails loudly if its CUDA
library is missing and never routes through here.

Parity pinning: ``oracle/make_golden.py`` lifts the reference's own classes out of
``/root/reference`` by ``ast`` (they cannot be imported: the scripts touch private paths
at import time), runs them on seeded inputs and stores the outputs in ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Reference locations restated here:
  * model            src/create_testset.py:40-83   (DoubleConv, UNet; logits = output of
                                                     final_conv, i.e. WITHOUT the :83 sigmoid)
  * BCE+Dice         train_bce_dice.py:186-199
  * focal / focal-Dice   src/train_with_focalDice.py:195-235
  * SDF / boundary / composite   src/train_with_boundary_loss.py:191-282
  * metrics          train_bce_dice.py:201-212, src/train_with_focalDice.py:266-284,
                     src/finetune_pseudo.py:192-208, src/finetune_for_224.py:223-233
  * (1,2,3)-dims dice variant   src/finetune_pseudo.py:178-190
  * pseudo-label threshold      src/data_preprocessing/create_pseudo_labels_gpu.py:201-215,294
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .edt_oracle import sdf_of_mask

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Model: state-dict layout of the reference UNet (src/create_testset.py:53-70)
# --------------------------------------------------------------------------------------
ENCODER = [("conv1", None, 64), ("conv2", 64, 128), ("conv3", 128, 256),
           ("conv4", 256, 512), ("conv5", 512, 1024)]
DECODER = [("upconv4", "dconv4", 1024, 512), ("upconv3", "dconv3", 512, 256),
           ("upconv2", "dconv2", 256, 128), ("upconv1", "dconv1", 128, 64)]


def state_dict_spec(in_channels: int = 3, out_channels: int = 1) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) for all 136 entries, in the order nn.Module.state_dict() yields them."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def double(prefix: str, cin: int, cout: int) -> None:
        for idx, ci in ((0, cin), (3, cout)):
            spec.append((f"{prefix}.conv.{idx}.weight", (cout, ci, 3, 3)))
            spec.append((f"{prefix}.conv.{idx}.bias", (cout,)))
            bn = idx + 1
            spec.append((f"{prefix}.conv.{bn}.weight", (cout,)))
            spec.append((f"{prefix}.conv.{bn}.bias", (cout,)))
            spec.append((f"{prefix}.conv.{bn}.running_mean", (cout,)))
            spec.append((f"{prefix}.conv.{bn}.running_var", (cout,)))
            spec.append((f"{prefix}.conv.{bn}.num_batches_tracked", ()))

    for name, cin, cout in ENCODER:
        double(name, in_channels if cin is None else cin, cout)
    for up, _, cin, cout in DECODER:
        spec.append((f"{up}.weight", (cin, cout, 2, 2)))
        spec.append((f"{up}.bias", (cout,)))
    for _, dc, cin, cout in DECODER:
        double(dc, cin, cout)
    spec.append(("final_conv.weight", (out_channels, 64, 1, 1)))
    spec.append(("final_conv.bias", (out_channels,)))
    return spec


def synth_state_dict(seed: int = 0, in_channels: int = 3, out_channels: int = 1,
                     dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic, RNG-library-independent weights (closed form), He-like magnitudes.

    Used for golden vectors so fixtures do not depend on torch's RNG stream.
    """
    sd: Dict[str, Tensor] = {}
    for i, (key, shape) in enumerate(state_dict_spec(in_channels, out_channels)):
        n = int(np.prod(shape)) if shape else 1
        idx = np.arange(n, dtype=np.float64)
        phase = 0.37 * (i + 1) + 0.011 * seed
        wave = np.sin(idx * (0.6180339887 + 0.001 * (i % 7)) + phase) \
            + 0.5 * np.cos(idx * 1.3247179572 + 2.0 * phase)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        if key.endswith("running_mean"):
            val = 0.05 * wave
        elif key.endswith("running_var"):
            val = 1.0 + 0.2 * np.abs(wave)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if key.startswith("upconv"):
                fan_in = shape[0] * 1  # ConvTranspose2d k2 s2: each output sees Cin inputs
            val = wave * math.sqrt(2.0 / fan_in) * 0.9
        elif ".conv.1." in key or ".conv.4." in key:        # BN affine
            val = (1.0 + 0.1 * wave) if key.endswith("weight") else 0.05 * wave
        else:                                                # conv / convT / head bias
            val = 0.02 * wave
        sd[key] = torch.from_numpy(val.reshape(shape)).to(dtype)
    return sd


def bf16_ste(t: Tensor) -> Tensor:
    """Round to bfloat16 (value), identity gradient (straight-through).  Used by the bf16-emulating variant
    of the model below, which places a rounding wherever the B200 kernels store a bf16 tensor — it separates
    "bf16 storage noise" from real defects when the fp32 oracle and the GPU path disagree."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def _double_conv(x: Tensor, sd: Dict[str, Tensor], p: str, training: bool, q=None, tape=None) -> Tensor:
    # src/create_testset.py:43-50 — Conv3x3(pad 1, bias) -> BN(eps 1e-5, momentum .1) -> ReLU, twice
    for c, b in ((0, 1), (3, 4)):
        w = sd[f"{p}.conv.{c}.weight"]
        x = F.conv2d(x, q(w) if q else w, sd[f"{p}.conv.{c}.bias"], padding=1)
        if q:
            # the kernels never add the conv bias before a train-mode BN (it cancels); in eval mode the bias is
            # folded into the epilogue affine, so only the train-mode pre-BN tensor is rounded
            x = q(x - sd[f"{p}.conv.{c}.bias"].view(1, -1, 1, 1)) + sd[f"{p}.conv.{c}.bias"].view(1, -1, 1, 1) \
                if training else x
        if tape is not None:
            tape[f"{p}.conv.{c}.y"] = x
            if x.requires_grad:
                x.retain_grad()
        x = F.batch_norm(x, sd[f"{p}.conv.{b}.running_mean"], sd[f"{p}.conv.{b}.running_var"],
                         sd[f"{p}.conv.{b}.weight"], sd[f"{p}.conv.{b}.bias"],
                         training=training, momentum=0.1, eps=1e-5)
        if training and f"{p}.conv.{b}.num_batches_tracked" in sd:
            sd[f"{p}.conv.{b}.num_batches_tracked"] += 1
        x = F.relu(x)
        if q:
            x = q(x)
        if tape is not None:
            tape[f"{p}.conv.{c}.out"] = x
            if x.requires_grad:
                x.retain_grad()
    return x


def unet_logits(x: Tensor, sd: Dict[str, Tensor], training: bool = False, emulate_bf16: bool = False,
                tape: Dict[str, Tensor] = None) -> Tensor:
    """src/create_testset.py:72-82 — returns the final_conv output (logits), sigmoid dropped.

    In training mode the BN running buffers inside ``sd`` are updated in place, as
    ``nn.BatchNorm2d`` does.  ``emulate_bf16`` rounds the input, the conv / conv-transpose weights and
    every stored activation to bfloat16 (straight-through gradients): the storage precision of the GPU
    path, everything else fp32.  ``tape`` (a dict) receives the per-layer tensors for stage-by-stage checks.
    """
    q = bf16_ste if emulate_bf16 else None
    skips = []
    h = q(x) if q else x
    for i, (name, _, _) in enumerate(ENCODER):
        if i:
            h = F.max_pool2d(h, 2, 2)
        h = _double_conv(h, sd, name, training, q, tape)
        skips.append(h)
    skips.pop()                                   # x5 is the bottleneck, not a skip
    for up, dc, _, _ in DECODER:
        w = sd[f"{up}.weight"]
        h = F.conv_transpose2d(h, q(w) if q else w, sd[f"{up}.bias"], stride=2)
        if q:
            h = q(h)
        if tape is not None:
            tape[f"{up}.out"] = h
            if h.requires_grad:
                h.retain_grad()
        h = torch.cat([h, skips.pop()], dim=1)    # upsampled first, then the skip (:78-81)
        h = _double_conv(h, sd, dc, training, q, tape)
    return F.conv2d(h, sd["final_conv.weight"], sd["final_conv.bias"])


def param_keys(sd: Dict[str, Tensor]) -> List[str]:
    """The 82 trainable tensors (everything except BN buffers)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


# --------------------------------------------------------------------------------------
# Losses
# --------------------------------------------------------------------------------------
def _dice_term(logits: Tensor, targets: Tensor, smooth: float, dims=(2, 3)) -> Tensor:
    p = torch.sigmoid(logits)
    inter = (p * targets).sum(dims)
    denom = p.sum(dims) + targets.sum(dims)
    return 1 - ((2 * inter + smooth) / (denom + smooth)).mean()


def bce_dice_loss(logits: Tensor, targets: Tensor, bce_weight: float = 0.5, smooth: float = 1.0,
                  dims=(2, 3)) -> Tensor:
    """train_bce_dice.py:193-199; dims=(1,2,3) gives src/finetune_pseudo.py:184-190."""
    bce = F.binary_cross_entropy_with_logits(logits, targets)
    return bce_weight * bce + (1 - bce_weight) * _dice_term(logits, targets, smooth, dims)


def bce_dice_loss_per_sample(logits: Tensor, targets: Tensor, bce_weight: float = 0.5,
                             smooth: float = 1.0) -> Tensor:
    """src/finetune_for_224.py:208-221 — returns [B]; note the hard-coded .5/.5 mix (:221)."""
    bce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none").mean(dim=(1, 2, 3))
    p = torch.sigmoid(logits)
    inter = (p * targets).sum(dim=(1, 2, 3))
    den = (p + targets).sum(dim=(1, 2, 3))
    return 0.5 * bce + 0.5 * (1 - (2 * inter + smooth) / (den + smooth))


def focal_loss(logits: Tensor, targets: Tensor, alpha: float = 0.25, gamma: float = 2.0,
               reduction: str = "mean") -> Tensor:
    """src/train_with_focalDice.py:207-219 — alpha is applied uniformly to both classes."""
    ce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    p = torch.sigmoid(logits)
    p_t = torch.where(targets == 1, p, 1 - p)
    out = alpha * (1 - p_t) ** gamma * ce
    if reduction == "mean":
        return out.mean()
    if reduction == "sum":
        return out.sum()
    return out


def focal_dice_loss(logits: Tensor, targets: Tensor, alpha: float = 0.5, gamma: float = 2.0,
                    smooth: float = 1.0, w_focal: float = 0.5) -> Tensor:
    """src/train_with_focalDice.py:229-235."""
    return w_focal * focal_loss(logits, targets, alpha, gamma) \
        + (1 - w_focal) * _dice_term(logits, targets, smooth)


@torch.no_grad()
def batch_sdf_from_masks(targets: Tensor) -> Tensor:
    """src/train_with_boundary_loss.py:204-217 — per image: >0.5, exact-EDT SDF, float32 / max(H,W)."""
    B, _, H, W = targets.shape
    out = np.empty((B, 1, H, W), dtype=np.float32)
    t = targets.detach().cpu().numpy()
    denom = np.float32(max(H, W))
    for b in range(B):
        sdf = sdf_of_mask(t[b, 0] > 0.5)                 # float32, sign: inside < 0 < outside
        out[b, 0] = sdf / denom                          # float32 / float32 (numpy-2 weak scalar)
    return torch.from_numpy(out).to(targets.device)


def symmetric_boundary_loss(logits: Tensor, targets: Tensor, t: float = 0.5, w_gt: float = 1.0,
                            w_pred: float = 0.5, use_abs: bool = True, scale: float = 1.0) -> Tensor:
    """src/train_with_boundary_loss.py:242-264."""
    p = torch.sigmoid(logits)
    sdf_gt = batch_sdf_from_masks(targets)
    with torch.no_grad():
        sdf_pred = batch_sdf_from_masks((p > t).float())
    a = p * sdf_gt
    b = (1.0 - p) * (-sdf_pred)
    if use_abs:
        a, b = a.abs(), b.abs()
    return scale * (w_gt * a.mean() + w_pred * b.mean())


def composite_seg_loss(logits: Tensor, targets: Tensor, bce_weight: float = 0.5,
                       boundary_weight: float = 0.3, sym_kwargs=None) -> Tensor:
    """src/train_with_boundary_loss.py:279-282."""
    reg = bce_dice_loss(logits, targets, bce_weight, 1.0)
    bnd = symmetric_boundary_loss(logits, targets, **(sym_kwargs or {}))
    return (1 - boundary_weight) * reg + boundary_weight * bnd


# --------------------------------------------------------------------------------------
# Metrics / thresholding
# --------------------------------------------------------------------------------------
@torch.no_grad()
def soft_dice_metric(logits: Tensor, targets: Tensor, smooth: float = 1.0, eps: float = 1e-7) -> float:
    """train_bce_dice.py:201-206."""
    p = torch.sigmoid(logits)
    inter = (p * targets).sum((2, 3))
    denom = p.sum((2, 3)) + targets.sum((2, 3))
    return ((2 * inter + smooth) / (denom + smooth + eps)).mean().item()


@torch.no_grad()
def hard_counts(logits: Tensor, targets: Tensor, t: float = 0.5, ge: bool = False):
    """Per-sample (pred_sum, target_sum, intersection) of the thresholded prediction.

    ``ge=False`` is ``sigmoid(x) > t`` (train_bce_dice.py:209); ``ge=True`` is ``>=``
    (create_pseudo_labels_gpu.py:294).
    """
    p = torch.sigmoid(logits)
    pred = ((p >= t) if ge else (p > t)).float()
    dims = tuple(range(1, logits.dim()))
    return pred.sum(dims), targets.sum(dims), (pred * targets).sum(dims)


@torch.no_grad()
def iou_metric(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> float:
    """train_bce_dice.py:208-212 (== src/finetune_pseudo.py:201-208 for C=1)."""
    ps, ts, inter = hard_counts(logits, targets, t)
    return ((inter + eps) / (ps + ts - inter + eps)).mean().item()


@torch.no_grad()
def hard_dice_metric(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7) -> float:
    """src/finetune_pseudo.py:192-199."""
    ps, ts, inter = hard_counts(logits, targets, t)
    return ((2 * inter + eps) / (ps + ts + eps)).mean().item()


@torch.no_grad()
def sweep_dice(logits: Tensor, targets: Tensor, thresholds: Iterable[float], smooth: float = 1.0) -> List[float]:
    """Inner expression of find_best_threshold, train_bce_dice.py:223-227, per threshold."""
    out = []
    for t in thresholds:
        ps, ts, inter = hard_counts(logits, targets, float(t))
        out.append(((2 * inter + smooth) / (ps + ts + smooth)).mean().item())
    return out


@torch.no_grad()
def precision_recall_f1(logits: Tensor, targets: Tensor, t: float = 0.5, eps: float = 1e-7):
    """src/train_with_focalDice.py:266-284."""
    pred = (torch.sigmoid(logits) > t).float()
    targets = targets.float()
    tp = (pred * targets).sum((2, 3))
    fp = (pred * (1 - targets)).sum((2, 3))
    fn = ((1 - pred) * targets).sum((2, 3))
    prec = tp / (tp + fp + eps)
    rec = tp / (tp + fn + eps)
    f1 = 2 * prec * rec / (prec + rec + eps)
    return tuple(torch.nan_to_num(v).mean().item() for v in (prec, rec, f1))


@torch.no_grad()
def pseudo_label_mask(logits: Tensor, threshold: float = 0.5) -> Tensor:
    """create_pseudo_labels_gpu.py:212,294 — sigmoid(logits)[:,0] >= threshold as uint8."""
    return (torch.sigmoid(logits)[:, 0] >= threshold).to(torch.uint8)


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def synth_batch(B: int, H: int, W: int, seed: int = 0, in_channels: int = 3):
    """Image ~ N(0,1); mask = one filled disc per sample (centre in the central half,
    radius in [H/11, H/3]).  Uses numpy's PCG64 so it is stable across torch versions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal((B, in_channels, H, W), dtype=np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((B, 1, H, W), dtype=np.float32)
    for b in range(B):
        cy = rng.uniform(H * 0.25, H * 0.75)
        cx = rng.uniform(W * 0.25, W * 0.75)
        r = rng.uniform(min(H, W) / 11.0, min(H, W) / 3.0)
        m[b, 0] = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(m)
