"""Drop-in ``nn.Module`` shells with the reference's names, constructor arguments and state-dict keys.

  UNet, DoubleConv                      src/create_testset.py:40-83
  BCEDiceLoss                           train_bce_dice.py:186-199   (dims=(1,2,3): src/finetune_pseudo.py:178-190)
  BCEDiceLossPerSample                  src/finetune_for_224.py:208-221
  FocalLoss, FocalDiceLoss              src/train_with_focalDice.py:195-235
  batch_sdf_from_masks, SymmetricBoundaryLoss, CompositeSegLoss
                                        src/train_with_boundary_loss.py:204-282

The modules own ordinary fp32 ``nn.Parameter``s (so optimizers, param groups, ``state_dict`` /
``load_state_dict`` and ``requires_grad`` toggling work unchanged); all arithmetic happens in the
``cartseg::`` ops.  Inputs must be CUDA tensors: there is no CPU path.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._lib import CartsegError


# =============================================================================================
# Model
# =============================================================================================
class DoubleConv(nn.Module):
    """Parameter container with the reference layout (conv.0 / conv.1 / conv.3 / conv.4).

    Only the whole-network op exists on the GPU, so calling a DoubleConv on its own raises."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, 3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, 3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):  # pragma: no cover - deliberate hard error
        raise CartsegError("DoubleConv is executed as part of cartseg.UNet (one fused op); it has no standalone path")


class _ParamGroup:
    """A view over some sub-modules of the UNet (``model.encoder`` etc.): lets training scripts build
    optimizer param groups and freeze / unfreeze them (src/train_with_focalDice.py:384-391,413-419)
    without registering duplicate state-dict entries."""

    def __init__(self, modules: Sequence[nn.Module]):
        self._modules_list = list(modules)

    def parameters(self) -> Iterator[nn.Parameter]:
        for m in self._modules_list:
            yield from m.parameters()

    def named_parameters(self):
        for i, m in enumerate(self._modules_list):
            for n, p in m.named_parameters():
                yield f"{i}.{n}", p

    def modules(self):
        for m in self._modules_list:
            yield from m.modules()

    def requires_grad_(self, flag: bool = True):
        for p in self.parameters():
            p.requires_grad_(flag)
        return self

    def train(self, mode: bool = True):
        """Per-group train / eval flags are accepted (scripts call ``model.encoder.eval()`` to freeze BN statistics),
        but the fused op has ONE BatchNorm mode: ``UNet.forward`` raises if the groups disagree with the model."""
        for m in self._modules_list:
            m.train(mode)
        return self

    def eval(self):
        return self.train(False)


_next_uid = 1


class UNet(nn.Module):
    """U-Net of src/create_testset.py:53-83 running as one B200-native op.

    ``forward`` returns **logits** (the output of ``final_conv``): every loss in the reference takes
    logits (train_bce_dice.py:194-195).  ``final_sigmoid=True`` reproduces the reference class's own
    ``torch.sigmoid`` tail (:83) for the interactive tool.
    """

    EVAL_CHUNK = 64          # eval-mode forwards of more images run in chunks of this many (measured best on B200)

    def __init__(self, in_channels: int = 3, out_channels: int = 1, final_sigmoid: bool = False):
        super().__init__()
        if out_channels != 1:
            raise CartsegError("cartseg.UNet implements the binary-segmentation head only (out_channels=1)")
        if not 1 <= in_channels <= 7:
            raise CartsegError("in_channels must be in [1, 7]")
        self.in_channels = in_channels
        self.final_sigmoid = final_sigmoid
        self.maxpool = nn.MaxPool2d(2, 2)
        self.conv1 = DoubleConv(in_channels, 64)
        self.conv2 = DoubleConv(64, 128)
        self.conv3 = DoubleConv(128, 256)
        self.conv4 = DoubleConv(256, 512)
        self.conv5 = DoubleConv(512, 1024)
        self.upconv4 = nn.ConvTranspose2d(1024, 512, 2, stride=2)
        self.upconv3 = nn.ConvTranspose2d(512, 256, 2, stride=2)
        self.upconv2 = nn.ConvTranspose2d(256, 128, 2, stride=2)
        self.upconv1 = nn.ConvTranspose2d(128, 64, 2, stride=2)
        self.dconv4 = DoubleConv(1024, 512)
        self.dconv3 = DoubleConv(512, 256)
        self.dconv2 = DoubleConv(256, 128)
        self.dconv1 = DoubleConv(128, 64)
        self.final_conv = nn.Conv2d(64, out_channels, 1)
        self._dp_handle = 0
        # bf16 weight packs follow (instance id, epoch): see ops._ensure_packed
        global _next_uid
        self._uid = _next_uid
        _next_uid += 1
        self._weights_epoch = 0          # advanced by every forward (eval too, unless freeze_packed) and load_state_dict
        self._packed_frozen = False
        self._reserve_sms = 0            # SMs left to the NCCL kernels in data-parallel runs (parallel.init_data_parallel)
        self.register_load_state_dict_post_hook(UNet._after_load_state_dict)

    @staticmethod
    def _after_load_state_dict(module, incompatible_keys) -> None:
        module._weights_epoch += 1

    def freeze_packed(self, flag: bool = True) -> "UNet":
        """Latency-critical inference (batch 1): promise that the parameters do not change between eval-mode forwards,
        so the bf16 weight packs are reused instead of being rebuilt on every call.  Training-mode forwards and
        ``load_state_dict`` still invalidate them; in-place edits through ``p.data`` do NOT — call this again (or
        ``freeze_packed(False)``) after such an edit."""
        self._packed_frozen = bool(flag)
        self._weights_epoch += 1
        return self

    def _check_single_bn_mode(self) -> None:
        mods = self._modules
        for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "dconv4", "dconv3", "dconv2", "dconv1"):
            seq = mods[name]._modules["conv"]._modules
            if seq["1"].training != self.training or seq["4"].training != self.training:
                raise CartsegError(
                    f"{name}: BatchNorm is in {'train' if seq['1'].training else 'eval'} mode but the model is in "
                    f"{'train' if self.training else 'eval'} mode.  cartseg.UNet runs as one fused op with a single "
                    "BatchNorm mode: per-group model.encoder.eval() (frozen BN statistics) is not supported")

    # ---- groups the training scripts address (smp.Unet naming) --------------------------------
    @property
    def encoder(self) -> _ParamGroup:
        return _ParamGroup([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5])

    @property
    def decoder(self) -> _ParamGroup:
        return _ParamGroup([self.upconv4, self.upconv3, self.upconv2, self.upconv1,
                            self.dconv4, self.dconv3, self.dconv2, self.dconv1])

    @property
    def segmentation_head(self) -> _ParamGroup:
        return _ParamGroup([self.final_conv])

    # ---- flat views in the order the C ABI expects (state-dict order) -------------------------
    def _double_convs(self) -> List[DoubleConv]:
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5,
                self.dconv4, self.dconv3, self.dconv2, self.dconv1]

    # The two flat views are rebuilt on every forward (parameters may have been re-assigned, moved or re-wrapped by the
    # caller), so they go through the modules' own dictionaries: nn.Module.__getattr__ costs ~1 us per access and made
    # these two functions 0.25 ms of host time per call — more than a batch-1 forward spends on the GPU.
    _DC_SLOTS = (("0", "weight"), ("0", "bias"), ("1", "weight"), ("1", "bias"),
                 ("3", "weight"), ("3", "bias"), ("4", "weight"), ("4", "bias"))

    def _flat_params(self) -> List[nn.Parameter]:
        mods = self._modules
        out: List[nn.Parameter] = []

        def dc(name: str):
            seq = mods[name]._modules["conv"]._modules
            for idx, pname in UNet._DC_SLOTS:
                out.append(seq[idx]._parameters[pname])

        for name in ("conv1", "conv2", "conv3", "conv4", "conv5"):
            dc(name)
        for name in ("upconv4", "upconv3", "upconv2", "upconv1"):
            pp = mods[name]._parameters
            out.append(pp["weight"])
            out.append(pp["bias"])
        for name in ("dconv4", "dconv3", "dconv2", "dconv1"):
            dc(name)
        pp = mods["final_conv"]._parameters
        out.append(pp["weight"])
        out.append(pp["bias"])
        return out

    def _flat_buffers(self) -> List[Tensor]:
        mods = self._modules
        out: List[Tensor] = []
        for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "dconv4", "dconv3", "dconv2", "dconv1"):
            seq = mods[name]._modules["conv"]._modules
            for idx in ("1", "4"):
                b = seq[idx]._buffers
                out.append(b["running_mean"])
                out.append(b["running_var"])
                out.append(b["num_batches_tracked"])
        return out

    def _frozen_encoder_convs(self, params: List[nn.Parameter]) -> int:
        n = 0
        for j in range(10):
            base = (j // 2) * 8 + (j % 2) * 4
            if any(params[i].requires_grad for i in range(base, base + 4)):
                break
            n += 1
        return n

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise CartsegError("cartseg.UNet takes CUDA tensors only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise CartsegError(f"expected input [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        B, C, H, W = x.shape
        if H % 16 or W % 16:
            raise CartsegError("H and W must be multiples of 16 (four 2x2 poolings)")
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        self._check_single_bn_mode()
        params = self._flat_params()
        buffers = self._flat_buffers()
        grad_mode = torch.is_grad_enabled()
        if (not self.training and not grad_mode and B > self.EVAL_CHUNK and not torch.compiler.is_compiling()):
            # Large eval batches run as chunks: per-sample independent (running statistics), and a chunk's activations
            # still overlap the 126 MB L2 between producer and consumer kernels — B = 512 in one piece measured 7 % slower
            # per image than B = 64 (14.9 k img/s at 64, 14.3 k at 128, 12.8 k at 512 in one piece).
            if not self._packed_frozen:
                self._weights_epoch += 1
            token = (self._uid << 40) | (self._weights_epoch & ((1 << 40) - 1))
            logits = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            for i in range(0, B, self.EVAL_CHUNK):
                n = min(self.EVAL_CHUNK, B - i)
                plan = ops.get_plan(n, C, H, W, x.device, inference_only=True)
                ops.unet_forward_impl(x[i:i + n], params, buffers, False, plan.id, token, out=logits[i:i + n])
            return torch.sigmoid(logits) if self.final_sigmoid else logits
        wants_grad = grad_mode and any(p.requires_grad for p in params)
        need_grad = self.training and wants_grad
        plan = ops.get_plan(B, C, H, W, x.device, inference_only=not self.training)
        if self.training or not self._packed_frozen:
            self._weights_epoch += 1
        token = (self._uid << 40) | (self._weights_epoch & ((1 << 40) - 1))
        if self._reserve_sms != plan.reserved_sms:
            ops.set_plan_sm_reserve(plan, self._reserve_sms)
        if need_grad:
            frozen = self._frozen_encoder_convs(params)
            logits = ops.UNetFunction.apply(x, plan.id, True, frozen, self._dp_handle, token,
                                            len(params), *params, *buffers)
        elif wants_grad and not torch.compiler.is_compiling():
            # eval mode with autograd on and trainable parameters: go through the autograd function so that a later
            # backward() gets the explicit error of UNetFunction.backward instead of a generic "does not require grad"
            logits = ops.UNetFunction.apply(x, plan.id, False, 0, 0, token, len(params), *params, *buffers)
        else:
            # without grad mode (torch.no_grad / inference_mode: every inference caller of the reference) the parameters
            # go to the op as they are; 82 detach() calls are 0.1 ms of host time
            if grad_mode or self.training or torch.compiler.is_compiling():
                plist = [p.detach() for p in params] if grad_mode else params
                logits = torch.ops.cartseg.unet_forward(x, plist, buffers, self.training, plan.id, token)
            else:                                        # eager inference: same body, without the dispatcher
                logits = ops.unet_forward_impl(x, params, buffers, False, plan.id, token)
        return torch.sigmoid(logits) if self.final_sigmoid else logits


# =============================================================================================
# Losses
# =============================================================================================
def _prep(logits: Tensor, targets: Tensor) -> Tuple[Tensor, Tensor]:
    if not logits.is_cuda:
        raise CartsegError("cartseg losses take CUDA tensors only (no CPU fallback)")
    if logits.shape != targets.shape:
        raise CartsegError(f"logits {tuple(logits.shape)} and targets {tuple(targets.shape)} differ in shape")
    lg = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.float().contiguous()
    tg = targets.detach()
    tg = tg if (tg.dtype == torch.float32 and tg.is_contiguous()) else tg.float().contiguous()
    return lg, tg


def _rows(logits: Tensor, dims: Tuple[int, ...]) -> int:
    if logits.dim() != 4:
        raise CartsegError("expected [B,C,H,W] logits")
    if tuple(dims) == (2, 3):
        return logits.shape[0] * logits.shape[1]
    if tuple(dims) == (1, 2, 3):
        return logits.shape[0]
    raise CartsegError("Dice dims must be (2,3) or (1,2,3)")


def _seg_loss(logits, targets, sdf_gt, sdf_pred, rows, *, w_elem=0.0, alpha=1.0, gamma=0.0, elem_sum=False,
              w_dice=0.0, smooth=1.0, w_bgt=0.0, w_bpred=0.0, use_abs=True, per_row=False) -> Tensor:
    loss, _ = torch.ops.cartseg.seg_loss(logits, targets, sdf_gt, sdf_pred, rows, float(w_elem), float(alpha),
                                         float(gamma), bool(elem_sum), float(w_dice), float(smooth), float(w_bgt),
                                         float(w_bpred), bool(use_abs), bool(per_row))
    return loss


class BCEDiceLoss(nn.Module):
    """train_bce_dice.py:186-199.  ``dims=(1,2,3)`` gives the variant of src/finetune_pseudo.py:178-190."""

    def __init__(self, bce_weight: float = 0.5, smooth: float = 1.0, dims: Tuple[int, ...] = (2, 3)):
        super().__init__()
        self.w, self.smooth, self.dims = bce_weight, smooth, tuple(dims)      # attribute names as in the reference

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        return _seg_loss(lg, tg, None, None, _rows(lg, self.dims), w_elem=self.w, alpha=1.0, gamma=0.0,
                         w_dice=1.0 - self.w, smooth=self.smooth)


class BCEDiceLossPerSample(nn.Module):
    """src/finetune_for_224.py:208-221 — returns one loss per sample ([B]); fixed .5/.5 mix (:221)."""

    def __init__(self, bce_weight: float = 0.5, smooth: float = 1.0):
        super().__init__()
        self.w, self.smooth = bce_weight, smooth

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        return _seg_loss(lg, tg, None, None, lg.shape[0], w_elem=0.5, alpha=1.0, gamma=0.0, w_dice=0.5,
                         smooth=self.smooth, per_row=True)


class FocalLoss(nn.Module):
    """src/train_with_focalDice.py:195-219 (alpha applied uniformly to both classes, :210-212)."""

    def __init__(self, alpha: float = 0.25, gamma: float = 2.0, reduction: str = "mean"):
        super().__init__()
        # 'mean' / 'sum': the fused reduction kernels; anything else returns the unreduced map, as the reference's
        # if / elif / else does (src/train_with_focalDice.py:214-219)
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        if self.reduction not in ("mean", "sum"):
            return torch.ops.cartseg.focal_map(lg, tg, float(self.alpha), float(self.gamma))
        return _seg_loss(lg, tg, None, None, lg.shape[0], w_elem=1.0, alpha=self.alpha, gamma=self.gamma,
                         elem_sum=self.reduction == "sum")


class FocalDiceLoss(nn.Module):
    """src/train_with_focalDice.py:221-235."""

    def __init__(self, alpha: float = 0.5, gamma: float = 2.0, smooth: float = 1.0, w_focal: float = 0.5):
        super().__init__()
        self.focal = FocalLoss(alpha=alpha, gamma=gamma, reduction="mean")
        self.smooth, self.w_focal = smooth, w_focal

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        return _seg_loss(lg, tg, None, None, _rows(lg, (2, 3)), w_elem=self.w_focal, alpha=self.focal.alpha,
                         gamma=self.focal.gamma, w_dice=1.0 - self.w_focal, smooth=self.smooth)


@torch.no_grad()
def batch_sdf_from_masks(targets: Tensor) -> Tensor:
    """src/train_with_boundary_loss.py:204-217 on the GPU: per image ``> 0.5`` -> exact-EDT signed
    distance (outside positive, inside negative, zeros for empty / full masks) -> ``/ max(H, W)``."""
    if targets.dim() != 4 or targets.shape[1] != 1:
        raise CartsegError("expected [B,1,H,W] masks")
    if not targets.is_cuda:
        raise CartsegError("cartseg.batch_sdf_from_masks takes CUDA tensors only (no CPU fallback)")
    t = targets.detach()
    t = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
    H, W = t.shape[2], t.shape[3]
    return torch.ops.cartseg.sdf(t, 0.5, False, float(max(H, W)))


def _pred_sdf(logits: Tensor, t: float) -> Tensor:
    # (sigmoid(x) > t).float() > 0.5  <=>  x >= x*(t)     (:250-251)
    H, W = logits.shape[2], logits.shape[3]
    return torch.ops.cartseg.sdf(logits.detach(), ops.logit_bound(t, ge=False), True, float(max(H, W)))


class SymmetricBoundaryLoss(nn.Module):
    """src/train_with_boundary_loss.py:219-264."""

    def __init__(self, t: float = 0.5, w_gt: float = 1.0, w_pred: float = 0.5, use_abs: bool = True,
                 scale: float = 1.0):
        super().__init__()
        self.t, self.w_gt, self.w_pred, self.use_abs, self.scale = t, w_gt, w_pred, use_abs, scale

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        if lg.shape[1] != 1:
            raise CartsegError("boundary loss expects single-channel logits")
        sdf_gt = batch_sdf_from_masks(tg)
        sdf_pred = _pred_sdf(lg, self.t)
        return _seg_loss(lg, tg, sdf_gt, sdf_pred, lg.shape[0], w_bgt=self.scale * self.w_gt,
                         w_bpred=self.scale * self.w_pred, use_abs=self.use_abs)


class CompositeSegLoss(nn.Module):
    """src/train_with_boundary_loss.py:266-282: (1-w)·BCEDice + w·SymmetricBoundary, one fused pass."""

    def __init__(self, bce_weight: float = 0.5, boundary_weight: float = 0.3, sym_kwargs: Optional[dict] = None):
        super().__init__()
        self.region = BCEDiceLoss(bce_weight=bce_weight, smooth=1.0)
        self.boundary = SymmetricBoundaryLoss(**(sym_kwargs or {}))
        self.wb = boundary_weight

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        lg, tg = _prep(logits, targets)
        if lg.shape[1] != 1:
            raise CartsegError("boundary loss expects single-channel logits")
        b, r, w = self.boundary, self.region, self.wb
        sdf_gt = batch_sdf_from_masks(tg)
        sdf_pred = _pred_sdf(lg, b.t)
        return _seg_loss(lg, tg, sdf_gt, sdf_pred, lg.shape[0],
                         w_elem=(1 - w) * r.w, alpha=1.0, gamma=0.0,
                         w_dice=(1 - w) * (1 - r.w), smooth=r.smooth,
                         w_bgt=w * b.scale * b.w_gt, w_bpred=w * b.scale * b.w_pred, use_abs=b.use_abs)


# =============================================================================================
# Active Boundary Loss ("next" row N1): src/training/losses/abl.py, src/training/train_BCEDice_ABL.py
# =============================================================================================
class ABL(nn.Module):
    """Drop-in for ``ABL`` (src/training/losses/abl.py:32-212), binary case ([B,1,H,W] logits, {0,1} targets).

    Same constructor arguments.  ``forward`` returns the loss tensor, or ``None`` when the predicted boundary is empty
    (abl.py:197-198) — deciding that costs one host read, exactly one where the reference has a dozen (its threshold
    loop, ``nonzero`` and per-image scipy EDTs).  ``forward_with_valid`` returns ``(loss, valid)`` without any host
    synchronisation; :class:`BCEDiceABL` uses it.  ``per_image_maps=True`` looks the GT distance up in image n's own
    map instead of reproducing the reference's batch indexing (behaviour 3).
    """

    def __init__(self, isdetach: bool = True, max_N_ratio: float = 1 / 100, ignore_label: int = 255,
                 label_smoothing: float = 0.2, weight=None, max_clip_dist: float = 20.0,
                 per_image_maps: bool = False):
        super().__init__()
        if not isdetach:
            raise NotImplementedError("ABL(isdetach=False): the reference never trains with attached neighbours")
        if weight is not None:
            raise NotImplementedError("ABL(weight=...) is only read by the label_smoothing == 0 branch of the reference")
        self.isdetach = isdetach
        self.max_N_ratio = max_N_ratio
        self.ignore_label = ignore_label
        self.label_smoothing = label_smoothing
        self.max_clip_dist = max_clip_dist
        self.per_image_maps = per_image_maps

    def forward_with_valid(self, logits: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
        if target.dim() == 3:
            target = target[:, None]
        logits, target = _prep(logits, target)
        loss, valid, _ = ops.abl_loss(logits, target, float(self.label_smoothing), float(self.max_N_ratio),
                                      float(self.max_clip_dist), int(self.ignore_label), bool(self.per_image_maps))
        return loss, valid

    def forward(self, logits: Tensor, target: Tensor) -> Optional[Tensor]:
        loss, valid = self.forward_with_valid(logits, target)
        return loss if bool(valid.item()) else None


class BCEDiceABL(nn.Module):
    """Drop-in for ``BCEDiceABL`` (src/training/train_BCEDice_ABL.py:264-302): region (BCE+Dice) + abl_weight * ABL,
    the region term alone when ABL has no valid boundary.  No host synchronisation in ``forward``; the
    ``boundary_none_count`` / ``total_calls`` counters of the reference are kept as device tensors."""

    def __init__(self, bce_weight: float = 0.5, smooth: float = 1.0, abl_weight: float = 0.1):
        super().__init__()
        self.region_loss = BCEDiceLoss(bce_weight=bce_weight, smooth=smooth)
        self.boundary_loss = ABL()
        self.abl_weight = abl_weight
        self.total_calls = 0
        self._none_count: Optional[Tensor] = None

    @property
    def boundary_none_count(self) -> int:
        return 0 if self._none_count is None else int(self._none_count.item())

    def components(self, logits: Tensor, targets: Tensor) -> dict:
        region = self.region_loss(logits, targets)
        boundary, valid = self.boundary_loss.forward_with_valid(logits, targets)
        self.total_calls += 1
        with torch.no_grad():
            miss = 1.0 - valid
            self._none_count = miss if self._none_count is None else self._none_count + miss
        # valid == 0: boundary is 0 with a zero gradient, so the sum is the region term alone
        total = region + self.abl_weight * boundary
        return {"total": total, "region": region.detach(), "boundary": boundary.detach()}

    def forward(self, logits: Tensor, targets: Tensor) -> Tensor:
        return self.components(logits, targets)["total"]
