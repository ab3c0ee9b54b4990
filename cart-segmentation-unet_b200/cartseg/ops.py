"""torch.library custom ops (namespace ``cartseg::``) over the C ABI of libcartseg.so.

Every op is CUDA-only; there is no CPU implementation and no PyTorch fallback.  The ops are thin:
argument checking, output allocation through torch (the caller owns every buffer, include/cartseg.h)
and one ctypes call that enqueues work on the current torch stream.

  cartseg::unet_forward / unet_backward      the whole U-Net (src/create_testset.py:40-83)
  cartseg::seg_loss / seg_loss_backward      BCE+Dice, focal(-Dice), boundary, composite losses
  cartseg::sdf                               exact-EDT signed distance maps
  cartseg::threshold_stats / threshold_mask  sigmoid -> threshold -> sums / uint8 masks
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import CartsegError, LossDesc, UnetTensors, check, ptr

# =============================================================================================
# U-Net plans: one per (batch, channels, H, W, device, inference_only); each owns its workspace.
# =============================================================================================
_MAX_PLANS = {False: 4, True: 6}          # training / inference plans kept alive per process (a ragged last batch, a
                                          # validation shape and the main shape must not evict each other every epoch)


class Plan:
    def __init__(self, plan_id: int, key, B, Cin, H, W, device, inference_only):
        self.id = plan_id
        self.key = key
        self.shape = (B, Cin, H, W)
        self.device = device
        self.inference_only = inference_only
        self.generation = 0               # bumped by every training-mode forward
        self.pack_key = None              # identity + version of the weights currently packed
        self.reserved_sms = 0
        self.handle = C.c_void_p()
        L = _lib.lib()
        check(L.cs_unet_plan_create(C.byref(self.handle), B, Cin, H, W, int(inference_only)), "cs_unet_plan_create")
        self.workspace_bytes = int(L.cs_unet_plan_workspace_bytes(self.handle))
        self.workspace = torch.empty(self.workspace_bytes + 1024, dtype=torch.uint8, device=device)
        base = self.workspace.data_ptr()
        aligned = (base + 1023) & ~1023
        with torch.cuda.device(device):
            check(L.cs_unet_plan_bind(self.handle, aligned, self.workspace_bytes), "cs_unet_plan_bind")

    def close(self):
        if self.handle:
            _lib.lib().cs_unet_plan_destroy(self.handle)
            self.handle = C.c_void_p()
        self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PLANS: Dict[int, Plan] = {}
_PLAN_BY_KEY: "OrderedDict[tuple, int]" = OrderedDict()
_next_plan_id = 1


def get_plan(B: int, Cin: int, H: int, W: int, device: torch.device, inference_only: bool) -> Plan:
    global _next_plan_id
    device = torch.device(device)
    if device.type != "cuda":
        raise CartsegError("cartseg U-Net runs on CUDA devices only (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (B, Cin, H, W, idx, bool(inference_only))
    pid = _PLAN_BY_KEY.get(key)
    if pid is not None:
        _PLAN_BY_KEY.move_to_end(key)
        return _PLANS[pid]
    # evict the least recently used plan of the same kind
    same = [k for k in _PLAN_BY_KEY if k[5] == bool(inference_only) and k[4] == idx]
    while len(same) >= _MAX_PLANS[bool(inference_only)]:
        old = same.pop(0)
        _PLANS.pop(_PLAN_BY_KEY.pop(old)).close()
    plan = Plan(_next_plan_id, key, B, Cin, H, W, torch.device("cuda", idx), bool(inference_only))
    _next_plan_id += 1
    _PLANS[plan.id] = plan
    _PLAN_BY_KEY[key] = plan.id
    return plan


def set_plan_sm_reserve(plan: Plan, reserve: int) -> None:
    """Leave ``reserve`` SMs to other resident kernels (NCCL) by capping the persistent GEMM grids."""
    n_sm = torch.cuda.get_device_properties(plan.device).multi_processor_count
    with torch.cuda.device(plan.device):
        check(_lib.lib().cs_unet_plan_set_sm_limit(plan.handle, n_sm - reserve if reserve > 0 else 0),
              "cs_unet_plan_set_sm_limit")
    plan.reserved_sms = reserve


def release_plans() -> None:
    """Drop every cached plan (and its workspace)."""
    for p in list(_PLANS.values()):
        p.close()
    _PLANS.clear()
    _PLAN_BY_KEY.clear()


def _plan(plan_id: int) -> Plan:
    p = _PLANS.get(plan_id)
    if p is None:
        raise CartsegError(f"U-Net plan {plan_id} no longer exists (evicted or released before its backward ran)")
    return p


_WEIGHT_SLOTS = tuple([b + o for b in list(range(0, 40, 8)) + list(range(48, 80, 8)) for o in (0, 4)]
                      + [40, 42, 44, 46])


def _fill_tensors(params: List[Tensor], grads: Optional[List[Optional[Tensor]]],
                  buffers: Optional[List[Tensor]]) -> UnetTensors:
    t = UnetTensors()
    if len(params) != _lib.NUM_PARAMS:
        raise CartsegError(f"expected {_lib.NUM_PARAMS} parameters in state-dict order, got {len(params)}")
    if buffers is not None and len(buffers) != 3 * _lib.NUM_BN:
        raise CartsegError(f"expected {3 * _lib.NUM_BN} BN buffers (mean, var, count per layer), got {len(buffers)}")
    # dtype / layout / device are checked on EVERY call (raw pointers go to the kernels): a cache keyed on object ids can
    # be hit by recycled ids after model.half() / .cpu() / a channels_last conversion.  ~20 us of host time.
    f32 = torch.float32
    for i, p in enumerate(params):
        if not p.is_cuda:
            raise CartsegError("cartseg ops take CUDA tensors only (no CPU fallback)")
        if p.dtype is not f32 or not p.is_contiguous():
            raise CartsegError(f"parameter {i} must be a contiguous float32 tensor")
    if buffers is not None:
        for i in range(_lib.NUM_BN):
            rm, rv, nbt = buffers[3 * i], buffers[3 * i + 1], buffers[3 * i + 2]
            if not (rm.is_cuda and rv.is_cuda and nbt.is_cuda):
                raise CartsegError("cartseg ops take CUDA tensors only (no CPU fallback)")
            if rm.dtype is not f32 or rv.dtype is not f32 or nbt.dtype is not torch.int64:
                raise CartsegError("BN buffers must be float32 running_mean / running_var and int64 num_batches_tracked")
    tp = t.param
    for i, p in enumerate(params):
        tp[i] = p.data_ptr()
    if grads is not None:
        tg = t.grad
        for i, g in enumerate(grads):
            tg[i] = ptr(g) if g is not None else None
    if buffers is not None:
        rm_, rv_, nb_ = t.running_mean, t.running_var, t.num_batches_tracked
        for i in range(_lib.NUM_BN):
            rm_[i] = buffers[3 * i].data_ptr()
            rv_[i] = buffers[3 * i + 1].data_ptr()
            nb_[i] = buffers[3 * i + 2].data_ptr()
    return t


def _ensure_packed(plan: Plan, params: List[Tensor], t: UnetTensors, pack_token: int) -> None:
    """Re-pack the bf16 weight copies when the fp32 parameters may have changed.  Tensor version counters are not
    enough (fused optimizers and ``p.data`` updates change parameters without bumping them; a freshly built model can
    reuse a freed model's addresses), so the module passes a token = (its unique instance id, a counter): the counter
    advances on every training-mode forward, every ``load_state_dict`` and — unless ``model.freeze_packed()`` was
    called — every eval-mode forward too.  The pack is one launch of ~30 us."""
    key = (pack_token,) + tuple((params[i].data_ptr(), params[i]._version) for i in _WEIGHT_SLOTS)
    if key != plan.pack_key:
        check(_lib.lib().cs_unet_pack_weights(plan.handle, C.byref(t), _lib.current_stream()), "cs_unet_pack_weights")
        plan.pack_key = key


def unet_forward_impl(x: Tensor, params: List[Tensor], buffers: List[Tensor], training: bool, plan_id: int,
                      pack_token: int, out: Optional[Tensor] = None) -> Tensor:
    """Body of ``cartseg::unet_forward``.  UNet.forward calls it directly for eager no-grad eval forwards: the
    dispatcher costs ~0.15 ms per call with 136 tensor arguments, as much as a batch-1 forward takes on the GPU."""
    plan = _plan(plan_id)
    B, Cin, H, W = plan.shape
    if tuple(x.shape) != (B, Cin, H, W) or x.dtype != torch.float32 or not x.is_contiguous():
        raise CartsegError(f"x must be a contiguous float32 tensor of shape {(B, Cin, H, W)}, got {tuple(x.shape)} {x.dtype}")
    if out is None:
        logits = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    else:                                   # a contiguous [B,1,H,W] slice of a larger output (chunked eval forwards)
        if tuple(out.shape) != (B, 1, H, W) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
            raise CartsegError("out must be a contiguous float32 [B,1,H,W] tensor on the input's device")
        logits = out
    t = _fill_tensors(params, None, buffers)
    with torch.cuda.device(x.device):
        _ensure_packed(plan, params, t, pack_token)
        check(_lib.lib().cs_unet_forward(plan.handle, C.byref(t), ptr(x), int(training), ptr(logits),
                                         _lib.current_stream()), "cs_unet_forward")
    if training:
        plan.generation += 1
    return logits


def _unet_forward_op(x: Tensor, params: List[Tensor], buffers: List[Tensor], training: bool, plan_id: int,
                     pack_token: int) -> Tensor:
    return unet_forward_impl(x, params, buffers, training, plan_id, pack_token)


unet_forward = torch.library.custom_op("cartseg::unet_forward", _unet_forward_op, mutates_args=("buffers",),
                                       device_types="cuda")


@unet_forward.register_fake
def _(x, params, buffers, training, plan_id, pack_token):
    return x.new_empty((x.shape[0], 1, x.shape[2], x.shape[3]), dtype=torch.float32)


# ---- backward ---------------------------------------------------------------------------------
_stage_params_cache: Optional[List[List[int]]] = None


def stage_params() -> List[List[int]]:
    """Parameter indices finalised by each of the 23 backward stages (reverse execution order)."""
    global _stage_params_cache
    if _stage_params_cache is None:
        L = _lib.lib()
        out = []
        buf = (C.c_int * 8)()
        for s in range(_lib.NUM_BWD_STAGES):
            n = L.cs_unet_stage_params(s, buf, 8)
            if n < 0:
                check(n, "cs_unet_stage_params")
            out.append([int(buf[i]) for i in range(n)])
        _stage_params_cache = out
    return _stage_params_cache


_DP_STATES: Dict[int, object] = {}         # handle -> parallel.GradSync (registered by cartseg.parallel)


def held_stages() -> Tuple[int, set]:
    """(flush stage, stages whose weight gradients are enqueued only when the flush stage is) — see
    cs_unet_backward_held_stages."""
    flush = C.c_int()
    buf = (C.c_int * _lib.NUM_BWD_STAGES)()
    n = _lib.lib().cs_unet_backward_held_stages(C.byref(flush), buf, _lib.NUM_BWD_STAGES)
    return int(flush.value), {int(buf[i]) for i in range(n)}


def grad_layout(params: List[Tensor]) -> Tuple[List[int], List[int], List[int]]:
    """Flat gradient buffer layout, in backward-stage order so that a data-parallel bucket of
    consecutive stages is one contiguous slice.  Returns (param order, element offset of each entry of
    that order, element offset at which each of the 23 stages starts (+ the total))."""
    order, offs, stage_off = [], [], []
    off = 0
    for st in stage_params():
        stage_off.append(off)
        for i in st:
            order.append(i)
            offs.append(off)
            off += params[i].numel()
    stage_off.append(off)
    return order, offs, stage_off


@torch.library.custom_op("cartseg::unet_backward", mutates_args=(), device_types="cuda")
def unet_backward(dlogits: Tensor, params: List[Tensor], plan_id: int, generation: int, frozen_encoder_convs: int,
                  dp_handle: int, no_grad_params: List[int]) -> Tensor:
    """Returns ONE flat float32 buffer holding every parameter gradient (layout: grad_layout).  Parameters listed in
    ``no_grad_params`` (requires_grad == False) get no gradient kernel at all — their weight-gradient GEMM is skipped —
    and their slice of the buffer is zero."""
    plan = _plan(plan_id)
    if generation != plan.generation:
        raise CartsegError("the activations of this forward pass were overwritten by a later training-mode forward "
                           "of the same shape; run backward before the next forward")
    B, _, H, W = plan.shape
    if tuple(dlogits.shape) != (B, 1, H, W):
        raise CartsegError(f"dlogits must have shape {(B, 1, H, W)}")
    dlogits = dlogits.to(torch.float32).contiguous()
    order, offs, stage_off = grad_layout(params)
    skip = set(int(i) for i in no_grad_params)
    for j in range(frozen_encoder_convs):              # a frozen encoder prefix is skipped by the kernels altogether
        base = (j // 2) * 8 + (j % 2) * 4
        skip.update(range(base, base + 4))
    # slices nobody writes must not be garbage (a data-parallel bucket all-reduces them)
    flat = (torch.zeros if skip else torch.empty)(stage_off[-1], dtype=torch.float32, device=dlogits.device)
    grads: List[Optional[Tensor]] = [None] * len(params)
    for i, o in zip(order, offs):
        if i not in skip:
            grads[i] = flat[o:o + params[i].numel()]
    t = _fill_tensors(params, grads, None)
    L = _lib.lib()
    sync = _DP_STATES.get(dp_handle) if dp_handle else None
    if dp_handle and sync is None:
        raise CartsegError(f"data-parallel handle {dp_handle} is not registered")
    with torch.cuda.device(dlogits.device):
        stream = _lib.current_stream()
        if sync is None:
            check(L.cs_unet_backward(plan.handle, C.byref(t), ptr(dlogits), 0, _lib.NUM_BWD_STAGES,
                                     frozen_encoder_convs, stream), "cs_unet_backward")
        else:
            # bucket by bucket; the join with the library's internal streams is deferred: only the communication
            # stream waits per bucket, the compute stream once at the end
            check(L.cs_unet_set_deferred_join(plan.handle, 1), "cs_unet_set_deferred_join")
            wait = lambda cuda_stream: check(L.cs_unet_backward_wait(plan.handle, cuda_stream), "cs_unet_backward_wait")  # noqa: E731
            flush_stage, held = held_stages()
            postponed: List[Tuple[int, int]] = []       # buckets with weight gradients that are not enqueued yet
            try:
                for (s0, s1) in sync.stage_buckets(stage_off):
                    check(L.cs_unet_backward(plan.handle, C.byref(t), ptr(dlogits), s0, s1, frozen_encoder_convs,
                                             stream), "cs_unet_backward")
                    flushed = s1 > flush_stage
                    if not flushed and any(st in held for st in range(s0, s1)):
                        postponed.append((s0, s1))
                        continue
                    if flushed:
                        for (p0, p1) in postponed:
                            sync.reduce_async(flat[stage_off[p0]:stage_off[p1]], wait=wait)
                        postponed.clear()
                    sync.reduce_async(flat[stage_off[s0]:stage_off[s1]], wait=wait)
                check(L.cs_unet_backward_wait(plan.handle, stream), "cs_unet_backward_wait")
            finally:
                L.cs_unet_set_deferred_join(plan.handle, 0)
            sync.finish()
    return flat


@unet_backward.register_fake
def _(dlogits, params, plan_id, generation, frozen_encoder_convs, dp_handle, no_grad_params):
    return dlogits.new_empty(sum(p.numel() for p in params), dtype=torch.float32)


class UNetFunction(torch.autograd.Function):
    """Autograd glue: forward / backward are the two cartseg:: ops above."""

    @staticmethod
    def forward(ctx, x, plan_id, training, frozen, dp_handle, pack_token, n_params, *tensors):
        params = list(tensors[:n_params])
        buffers = list(tensors[n_params:])
        logits = torch.ops.cartseg.unet_forward(x, [p.detach() for p in params], buffers, training, plan_id,
                                                pack_token)
        ctx.plan_id = plan_id
        ctx.x_version = x._version
        ctx.generation = _plan(plan_id).generation
        ctx.training = training
        ctx.frozen = frozen
        ctx.dp_handle = dp_handle
        ctx.n_params = n_params
        ctx.n_buffers = len(buffers)
        # x is saved too: the backward pass re-reads the image for the first convolution's weight gradient
        # (include/cartseg.h, cs_unet_forward) — this only keeps the caller's tensor alive, nothing is copied
        ctx.save_for_backward(x, *params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if not ctx.training:
            raise CartsegError("backward through an eval-mode forward is not available: call model.train() "
                               "(batch-statistics BN) for training steps")
        x = ctx.saved_tensors[0]
        if x._version != ctx.x_version:
            raise CartsegError("the input image was modified in place between forward and backward; the backward pass "
                               "re-reads it for the first convolution's weight gradient")
        params = [p.detach() for p in ctx.saved_tensors[1:]]
        need = ctx.needs_input_grad[7:7 + ctx.n_params]
        flat = torch.ops.cartseg.unet_backward(dlogits, params, ctx.plan_id, ctx.generation, ctx.frozen, ctx.dp_handle,
                                               [i for i in range(ctx.n_params) if not need[i]])
        order, offs, _ = grad_layout(params)
        out: List[Optional[Tensor]] = [None] * ctx.n_params
        for i, o in zip(order, offs):
            if need[i]:
                out[i] = flat[o:o + params[i].numel()].view(params[i].shape)
        return (None, None, None, None, None, None, None, *out, *([None] * ctx.n_buffers))


# =============================================================================================
# Losses
# =============================================================================================
def _desc(rows: int, n: int, w_elem: float, alpha: float, gamma: float, elem_sum: bool, w_dice: float, smooth: float,
          w_bgt: float, w_bpred: float, use_abs: bool, per_row: bool) -> LossDesc:
    return LossDesc(rows, n, w_elem, alpha, gamma, int(elem_sum), w_dice, smooth, w_bgt, w_bpred, int(use_abs),
                    int(per_row))


def _check_loss_inputs(logits: Tensor, targets: Tensor, sdf_gt: Optional[Tensor], sdf_pred: Optional[Tensor],
                       rows: int) -> int:
    for name, v in (("logits", logits), ("targets", targets), ("sdf_gt", sdf_gt), ("sdf_pred", sdf_pred)):
        if v is None:
            continue
        if not v.is_cuda:
            raise CartsegError(f"{name} must be a CUDA tensor (no CPU fallback)")
        if v.dtype != torch.float32 or not v.is_contiguous() or v.numel() != logits.numel():
            raise CartsegError(f"{name} must be contiguous float32 with {logits.numel()} elements")
    if rows < 1 or logits.numel() % rows:
        raise CartsegError("rows must divide the number of elements")
    n = logits.numel() // rows
    if n % 4:
        raise CartsegError("elements per Dice row must be a multiple of 4")
    return n


@torch.library.custom_op("cartseg::seg_loss", mutates_args=(), device_types="cuda")
def seg_loss(logits: Tensor, targets: Tensor, sdf_gt: Optional[Tensor], sdf_pred: Optional[Tensor], rows: int,
             w_elem: float, alpha: float, gamma: float, elem_sum: bool, w_dice: float, smooth: float, w_bgt: float,
             w_bpred: float, use_abs: bool, per_row: bool) -> Tuple[Tensor, Tensor]:
    n = _check_loss_inputs(logits, targets, sdf_gt, sdf_pred, rows)
    L = _lib.lib()
    scratch = torch.empty(int(L.cs_loss_scratch_bytes(rows)) // 8 + 1, dtype=torch.float64, device=logits.device)
    out = torch.empty(rows if per_row else 1, dtype=torch.float32, device=logits.device)
    d = _desc(rows, n, w_elem, alpha, gamma, elem_sum, w_dice, smooth, w_bgt, w_bpred, use_abs, per_row)
    with torch.cuda.device(logits.device):
        check(L.cs_loss_forward(C.byref(d), ptr(logits), ptr(targets), ptr(sdf_gt), ptr(sdf_pred), ptr(scratch),
                                ptr(out), _lib.current_stream()), "cs_loss_forward")
    return (out if per_row else out.reshape(())), scratch


@seg_loss.register_fake
def _(logits, targets, sdf_gt, sdf_pred, rows, w_elem, alpha, gamma, elem_sum, w_dice, smooth, w_bgt, w_bpred,
      use_abs, per_row):
    out = logits.new_empty((rows,) if per_row else (), dtype=torch.float32)
    return out, logits.new_empty(rows * 8 + 3, dtype=torch.float64)


@torch.library.custom_op("cartseg::seg_loss_backward", mutates_args=(), device_types="cuda")
def seg_loss_backward(grad_out: Tensor, logits: Tensor, targets: Tensor, sdf_gt: Optional[Tensor],
                      sdf_pred: Optional[Tensor], scratch: Tensor, rows: int, w_elem: float, alpha: float,
                      gamma: float, elem_sum: bool, w_dice: float, smooth: float, w_bgt: float, w_bpred: float,
                      use_abs: bool, per_row: bool) -> Tensor:
    n = _check_loss_inputs(logits, targets, sdf_gt, sdf_pred, rows)
    go = grad_out.to(torch.float32).contiguous().reshape(-1)
    if go.numel() != (rows if per_row else 1):
        raise CartsegError("grad_out has the wrong number of elements")
    dlogits = torch.empty_like(logits)
    d = _desc(rows, n, w_elem, alpha, gamma, elem_sum, w_dice, smooth, w_bgt, w_bpred, use_abs, per_row)
    with torch.cuda.device(logits.device):
        check(_lib.lib().cs_loss_backward(C.byref(d), ptr(logits), ptr(targets), ptr(sdf_gt), ptr(sdf_pred),
                                          ptr(scratch), ptr(go), ptr(dlogits), _lib.current_stream()),
              "cs_loss_backward")
    return dlogits


@seg_loss_backward.register_fake
def _(grad_out, logits, *args):
    return torch.empty_like(logits)


def _seg_loss_setup(ctx, inputs, output):
    logits, targets, sdf_gt, sdf_pred = inputs[:4]
    ctx.scalars = inputs[4:]
    ctx.has = (sdf_gt is not None, sdf_pred is not None)
    saved = [logits, targets] + [v for v in (sdf_gt, sdf_pred) if v is not None] + [output[1]]
    ctx.save_for_backward(*saved)


def _seg_loss_bwd(ctx, grad_loss, grad_scratch):
    saved = list(ctx.saved_tensors)
    logits, targets = saved[0], saved[1]
    k = 2
    sdf_gt = sdf_pred = None
    if ctx.has[0]:
        sdf_gt = saved[k]; k += 1
    if ctx.has[1]:
        sdf_pred = saved[k]; k += 1
    scratch = saved[k]
    dlogits = torch.ops.cartseg.seg_loss_backward(grad_loss, logits, targets, sdf_gt, sdf_pred, scratch, *ctx.scalars)
    return (dlogits,) + (None,) * 14


seg_loss.register_autograd(_seg_loss_bwd, setup_context=_seg_loss_setup)
# under torch.autocast the loss takes float32 like the reference's native op (label_smooth.py:63 custom_fwd(cast_inputs=float32))
seg_loss.register_autocast("cuda", torch.float32)


# ---- FocalLoss(reduction="none"): the unreduced map (src/train_with_focalDice.py:214-219) --------------------------------
@torch.library.custom_op("cartseg::focal_map", mutates_args=(), device_types="cuda")
def focal_map(logits: Tensor, targets: Tensor, alpha: float, gamma: float) -> Tensor:
    for name, v in (("logits", logits), ("targets", targets)):
        if not v.is_cuda:
            raise CartsegError(f"cartseg::focal_map: {name} must be a CUDA tensor (no CPU fallback)")
        if v.dtype != torch.float32 or not v.is_contiguous() or v.numel() != logits.numel():
            raise CartsegError(f"cartseg::focal_map: {name} must be contiguous float32 with {logits.numel()} elements")
    out = torch.empty_like(logits)
    if logits.numel():
        with torch.cuda.device(logits.device):
            check(_lib.lib().cs_focal_map_forward(ptr(logits), ptr(targets), logits.numel(), alpha, gamma, ptr(out),
                                                  _lib.current_stream()), "cs_focal_map_forward")
    return out


@focal_map.register_fake
def _(logits, targets, alpha, gamma):
    return torch.empty_like(logits)


@torch.library.custom_op("cartseg::focal_map_backward", mutates_args=(), device_types="cuda")
def focal_map_backward(grad_out: Tensor, logits: Tensor, targets: Tensor, alpha: float, gamma: float) -> Tensor:
    go = grad_out.to(torch.float32).contiguous()
    dlogits = torch.empty_like(logits)
    if logits.numel():
        with torch.cuda.device(logits.device):
            check(_lib.lib().cs_focal_map_backward(ptr(logits), ptr(targets), ptr(go), logits.numel(), alpha, gamma,
                                                   ptr(dlogits), _lib.current_stream()), "cs_focal_map_backward")
    return dlogits


@focal_map_backward.register_fake
def _(grad_out, logits, targets, alpha, gamma):
    return torch.empty_like(logits)


def _focal_map_setup(ctx, inputs, output):
    ctx.scalars = inputs[2:]
    ctx.save_for_backward(inputs[0], inputs[1])


def _focal_map_bwd(ctx, grad):
    logits, targets = ctx.saved_tensors
    return torch.ops.cartseg.focal_map_backward(grad, logits, targets, *ctx.scalars), None, None, None


focal_map.register_autograd(_focal_map_bwd, setup_context=_focal_map_setup)
focal_map.register_autocast("cuda", torch.float32)


# =============================================================================================
# Active Boundary Loss (src/training/losses/abl.py:66-212)
# =============================================================================================
def abl_eps_ladder() -> List[float]:
    """The thresholds the reference's ``while True: eps *= 1.2`` loop (abl.py:78-83) visits, computed the way Python
    does (float64 products); the C side rounds them to float32, which is what the tensor comparison sees."""
    out, e = [], 1e-5
    for _ in range(_lib.ABL_LADDER):
        out.append(e)
        e *= 1.2
    return out


def _abl_desc(B: int, H: int, W: int, label_smoothing: float, max_n_ratio: float, max_clip_dist: float,
              ignore_label: int, per_image_maps: bool) -> "_lib.AblDesc":
    import numpy as np
    d = _lib.AblDesc()
    d.batch, d.height, d.width = B, H, W
    d.max_n = float(np.float32((H * W) * max_n_ratio))        # abl.py:69; compared in float32 by torch
    d.label_smoothing, d.max_clip_dist = label_smoothing, max_clip_dist
    d.ignore_label, d.per_image_maps = ignore_label, int(per_image_maps)
    for i, e in enumerate(abl_eps_ladder()):
        d.eps_ladder[i] = e
    return d


def _check_abl_inputs(logits: Tensor, targets: Tensor) -> Tuple[int, int, int]:
    if not (logits.is_cuda and targets.is_cuda):
        raise CartsegError("cartseg::abl_loss takes CUDA tensors only (no CPU fallback)")
    if logits.dim() != 4 or logits.shape[1] != 1:
        raise CartsegError("abl_loss: logits must be [B,1,H,W] (the binary case the reference trains)")
    if logits.dtype != torch.float32 or targets.dtype != torch.float32:
        raise CartsegError("abl_loss: logits and targets must be float32")
    if not (logits.is_contiguous() and targets.is_contiguous()) or targets.numel() != logits.numel():
        raise CartsegError("abl_loss: logits / targets must be contiguous and of the same size")
    return logits.shape[0], logits.shape[2], logits.shape[3]


def _abl_scratch(B: int, H: int, W: int, device) -> Tensor:
    n = int(_lib.lib().cs_abl_scratch_bytes(B, H, W))
    return torch.empty(n, dtype=torch.uint8, device=device)       # torch's caching allocator aligns to 512 B


@torch.library.custom_op("cartseg::abl_loss", mutates_args=(), device_types="cuda")
def abl_loss(logits: Tensor, targets: Tensor, label_smoothing: float, max_n_ratio: float, max_clip_dist: float,
             ignore_label: int, per_image_maps: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (loss [], valid [] (1.0 / 0.0: the reference returns None when 0), scratch)."""
    B, H, W = _check_abl_inputs(logits, targets)
    scratch = _abl_scratch(B, H, W, logits.device)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    d = _abl_desc(B, H, W, label_smoothing, max_n_ratio, max_clip_dist, ignore_label, per_image_maps)
    with torch.cuda.device(logits.device):
        check(_lib.lib().cs_abl_forward(C.byref(d), ptr(logits), ptr(targets), ptr(scratch), ptr(out),
                                        _lib.current_stream()), "cs_abl_forward")
    return out[0].clone(), out[1].clone(), scratch


@abl_loss.register_fake
def _(logits, targets, label_smoothing, max_n_ratio, max_clip_dist, ignore_label, per_image_maps):
    B, H, W = logits.shape[0], logits.shape[2], logits.shape[3]
    return (logits.new_empty(()), logits.new_empty(()),
            logits.new_empty(B * H * W * 10 + 4096, dtype=torch.uint8))


@torch.library.custom_op("cartseg::abl_loss_backward", mutates_args=(), device_types="cuda")
def abl_loss_backward(grad_out: Tensor, logits: Tensor, scratch: Tensor, label_smoothing: float, max_n_ratio: float,
                      max_clip_dist: float, ignore_label: int, per_image_maps: bool) -> Tensor:
    B, H, W = logits.shape[0], logits.shape[2], logits.shape[3]
    go = grad_out.to(torch.float32).contiguous().reshape(-1)
    dlogits = torch.empty_like(logits)
    d = _abl_desc(B, H, W, label_smoothing, max_n_ratio, max_clip_dist, ignore_label, per_image_maps)
    with torch.cuda.device(logits.device):
        check(_lib.lib().cs_abl_backward(C.byref(d), ptr(logits), ptr(scratch), ptr(go), ptr(dlogits),
                                         _lib.current_stream()), "cs_abl_backward")
    return dlogits


@abl_loss_backward.register_fake
def _(grad_out, logits, *args):
    return torch.empty_like(logits)


def _abl_setup(ctx, inputs, output):
    ctx.scalars = inputs[2:]
    ctx.save_for_backward(inputs[0], output[2])


def _abl_bwd(ctx, grad_loss, grad_valid, grad_scratch):
    logits, scratch = ctx.saved_tensors
    return (torch.ops.cartseg.abl_loss_backward(grad_loss, logits, scratch, *ctx.scalars),) + (None,) * 6


abl_loss.register_autograd(_abl_bwd, setup_context=_abl_setup)
abl_loss.register_autocast("cuda", torch.float32)


def abl_debug(logits: Tensor, scratch: Tensor, per_image_maps: bool = False):
    """Test hook: (eps, ladder index, kept pixels, predicted-boundary pixels, distance map [B,H,W] uint16 (numpy),
    kl map [B,H,W] float32 (numpy)) of the forward call that filled ``scratch``.  Synchronises."""
    import numpy as np
    B, H, W = logits.shape[0], logits.shape[2], logits.shape[3]
    d = _abl_desc(B, H, W, 0.2, 0.01, 20.0, 255, per_image_maps)
    eps, k = C.c_float(), C.c_int()
    kept, nb = C.c_ulonglong(), C.c_ulonglong()
    dm = np.empty((B, H, W), np.uint16)
    kl = np.empty((B, H, W), np.float32)
    with torch.cuda.device(logits.device):
        check(_lib.lib().cs_abl_debug_read(C.byref(d), ptr(scratch), C.byref(eps), C.byref(k), C.byref(kept),
                                           C.byref(nb), dm.ctypes.data, kl.ctypes.data, _lib.current_stream()),
              "cs_abl_debug_read")
    return eps.value, k.value, kept.value, nb.value, dm, kl


# =============================================================================================
# Signed distance maps
# =============================================================================================
@torch.library.custom_op("cartseg::sdf", mutates_args=(), device_types="cuda")
def sdf(src: Tensor, thr: float, ge: bool, norm: float) -> Tensor:
    """src [B,1,H,W] or [B,H,W] float32.  fg = src >= thr if ge else src > thr."""
    if not src.is_cuda:
        raise CartsegError("cartseg::sdf takes CUDA tensors only (no CPU fallback)")
    if src.dtype != torch.float32 or not src.is_contiguous():
        raise CartsegError("src must be contiguous float32")
    if src.dim() == 4 and src.shape[1] == 1:
        B, H, W = src.shape[0], src.shape[2], src.shape[3]
    elif src.dim() == 3:
        B, H, W = src.shape
    else:
        raise CartsegError("src must be [B,1,H,W] or [B,H,W]")
    out = torch.empty_like(src)
    if src.numel() == 0:
        return out
    L = _lib.lib()
    scratch = torch.empty(int(L.cs_sdf_scratch_bytes(B, H, W)), dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        check(L.cs_sdf(ptr(src), thr, int(ge), B, H, W, norm, ptr(out), ptr(scratch), _lib.current_stream()), "cs_sdf")
    return out


@sdf.register_fake
def _(src, thr, ge, norm):
    return torch.empty_like(src)


# =============================================================================================
# Thresholding / metric sums
# =============================================================================================
@torch.library.custom_op("cartseg::threshold_stats", mutates_args=(), device_types="cuda")
def threshold_stats(logits: Tensor, targets: Tensor, rows: int, xs: Tensor) -> Tuple[Tensor, Tensor]:
    """counts [rows, K, 2] = (sum pred, sum pred*t) with pred = logits >= xs[k];
    soft [rows, 3] = (sum p, sum t, sum p*t).  float64."""
    for v in (logits, targets, xs):
        if not v.is_cuda or v.dtype != torch.float32 or not v.is_contiguous():
            raise CartsegError("threshold_stats takes contiguous float32 CUDA tensors")
    if logits.numel() != targets.numel() or rows < 1 or logits.numel() % rows:
        raise CartsegError("logits / targets size mismatch")
    K = xs.numel()
    n = logits.numel() // rows
    counts = torch.empty((rows, K, 2), dtype=torch.float64, device=logits.device)
    soft = torch.empty((rows, 3), dtype=torch.float64, device=logits.device)
    L = _lib.lib()
    with torch.cuda.device(logits.device):
        stream = _lib.current_stream()
        for k0 in range(0, K, 32):                   # the kernel handles up to 32 thresholds per pass
            k1 = min(K, k0 + 32)
            part = counts if K <= 32 else torch.empty((rows, k1 - k0, 2), dtype=torch.float64, device=logits.device)
            check(L.cs_threshold_stats(ptr(logits), ptr(targets), rows, n, xs.data_ptr() + 4 * k0, k1 - k0, ptr(part),
                                       ptr(soft), stream), "cs_threshold_stats")
            if part is not counts:
                counts[:, k0:k1] = part
    return counts, soft


@threshold_stats.register_fake
def _(logits, targets, rows, xs):
    return (logits.new_empty((rows, xs.numel(), 2), dtype=torch.float64),
            logits.new_empty((rows, 3), dtype=torch.float64))


@torch.library.custom_op("cartseg::threshold_mask", mutates_args=(), device_types="cuda")
def threshold_mask(logits: Tensor, xstar: float) -> Tensor:
    """uint8 mask = logits >= xstar (same shape as logits)."""
    if not logits.is_cuda or logits.dtype != torch.float32 or not logits.is_contiguous():
        raise CartsegError("threshold_mask takes a contiguous float32 CUDA tensor")
    if logits.numel() % 4:
        raise CartsegError("number of logits must be a multiple of 4")
    mask = torch.empty(logits.shape, dtype=torch.uint8, device=logits.device)
    if logits.numel():
        with torch.cuda.device(logits.device):
            check(_lib.lib().cs_threshold_mask(ptr(logits), logits.numel(), xstar, ptr(mask), _lib.current_stream()),
                  "cs_threshold_mask")
    return mask


@threshold_mask.register_fake
def _(logits, xstar):
    return logits.new_empty(logits.shape, dtype=torch.uint8)


# ---------------------------------------------------------------------------------------------
# sigmoid(x) > t  <=>  x >= x*(t): the smallest float32 whose sigmoid passes the comparison.
# Found by bisection over float32 bit patterns with torch's own float32 sigmoid, so thresholded
# masks are bit-identical to the reference's sigmoid-then-compare (train_bce_dice.py:209 uses >,
# create_pseudo_labels_gpu.py:294 uses >=) without materialising probabilities.
# ---------------------------------------------------------------------------------------------
_bound_cache: Dict[Tuple[float, bool], float] = {}


def _f32_from_ordered(u: int) -> float:
    import struct
    # ordered-int -> float32: monotone map of int32 order onto float order
    if u < 0:
        u = -(u + 1)
        bits = (u & 0x7FFFFFFF) | 0x80000000
    else:
        bits = u
    return struct.unpack("<f", struct.pack("<I", bits & 0xFFFFFFFF))[0]


def logit_bound(t: float, ge: bool = False) -> float:
    """Smallest float32 x with sigmoid(x) > t (or >= t if ge); +inf if none."""
    key = (float(t), bool(ge))
    if key in _bound_cache:
        return _bound_cache[key]
    tt = torch.tensor(float(t), dtype=torch.float32)

    def passes(x: float) -> bool:
        # a 16-wide tensor so that ATen takes its vectorised path, as it does for whole logit maps
        p = torch.sigmoid(torch.full((16,), x, dtype=torch.float32))[0]
        return bool(p >= tt) if ge else bool(p > tt)

    lo, hi = -0x7F7FFFFF - 1, 0x7F7FFFFF          # ordered ints of -FLT_MAX .. +FLT_MAX
    if passes(_f32_from_ordered(lo)):
        res = float("-inf")
    elif not passes(_f32_from_ordered(hi)):
        res = float("inf")
    else:
        while hi - lo > 1:                          # invariant: !passes(lo), passes(hi)
            mid = (lo + hi) // 2
            if passes(_f32_from_ordered(mid)):
                hi = mid
            else:
                lo = mid
        res = _f32_from_ordered(hi)
    _bound_cache[key] = res
    return res
