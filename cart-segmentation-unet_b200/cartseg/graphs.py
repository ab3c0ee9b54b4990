"""CUDA-graph wrappers of the hot path ("no host sync, CUDA-graph capturable").

Every ``cs_*`` entry point only enqueues work on the caller's stream (the backward pass forks / joins its two internal
streams with events created when the plan is bound), so a whole training step — forward, loss, backward — or an
eval-mode forward + threshold can be captured once and replayed with ONE launch.  That matters where the step is
launch-bound: batch-1 ... 8 inference (30 kernels of a few microseconds each; the pseudo-label generator of
src/data_preprocessing/create_pseudo_labels_gpu.py:266-294 at small batch) and small-image fine-tuning.

Capture follows torch's whole-network recipe (static input / output tensors, a few warm-up iterations on a side
stream first so that plans, workspaces and function attributes exist before the capture starts).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import Tensor

from . import ops
from ._lib import CartsegError


def _warmup(fn: Callable[[], None], iters: int) -> None:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(iters):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()


class GraphedInference:
    """``masks_or_logits = GraphedInference(model, example_x, threshold=0.5)(x)``

    Eval-mode forward of a :class:`cartseg.UNet` (+ ``sigmoid(logits) >= threshold`` as a uint8 mask when a threshold
    is given — create_pseudo_labels_gpu.py:294) replayed from a CUDA graph.  The returned tensor is the graph's static
    output: it is overwritten by the next call.  Unless ``model.freeze_packed()`` was called, the bf16 weight re-pack is
    part of the graph, so parameter updates made between calls are seen."""

    def __init__(self, model, example_x: Tensor, threshold: Optional[float] = None, warmup: int = 3):
        if not example_x.is_cuda:
            raise CartsegError("GraphedInference takes CUDA tensors only (no CPU fallback)")
        if model.training:
            raise CartsegError("GraphedInference captures an eval-mode forward: call model.eval() first")
        self.model = model
        self.x = example_x.detach().to(torch.float32).contiguous().clone()
        self.xstar = None if threshold is None else ops.logit_bound(float(threshold), ge=True)

        def fwd():
            with torch.no_grad():
                z = model(self.x)
                return z if self.xstar is None else torch.ops.cartseg.threshold_mask(z, self.xstar)

        _warmup(fwd, warmup)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fwd()

    def __call__(self, x: Tensor) -> Tensor:
        if x.shape != self.x.shape:
            raise CartsegError(f"graph captured for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out


class GraphedTrainStep:
    """Forward + criterion + backward of one (batch, H, W) captured in a CUDA graph.

    ``loss = step(x, target)`` copies the batch into the static inputs, replays the graph and returns the static loss
    tensor; parameter gradients land in the parameters' (static) ``.grad`` tensors, so an ordinary
    ``optimizer.step()`` follows.  Do not call ``optimizer.zero_grad(set_to_none=True)`` between steps — the graph
    overwrites the gradients in place."""

    def __init__(self, model, criterion, example_x: Tensor, example_t: Tensor, warmup: int = 3):
        if not (example_x.is_cuda and example_t.is_cuda):
            raise CartsegError("GraphedTrainStep takes CUDA tensors only (no CPU fallback)")
        if not model.training:
            raise CartsegError("GraphedTrainStep captures a training-mode step: call model.train() first")
        if getattr(model, "_dp_handle", 0):
            raise CartsegError("GraphedTrainStep: the data-parallel all-reduce is not captured; use the eager step")
        self.model, self.criterion = model, criterion
        self.x = example_x.detach().to(torch.float32).contiguous().clone()
        self.t = example_t.detach().to(torch.float32).contiguous().clone()

        def step():
            for p in model.parameters():
                p.grad = None
            loss = criterion(model(self.x), self.t)
            loss.backward()
            return loss

        # warm-up steps DO update the BN running statistics (as any training step does); parameters are untouched
        _warmup(step, warmup)
        for p in model.parameters():
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss = criterion(model(self.x), self.t)
            loss.backward()
            self.loss = loss.detach()
        self.static_grads = [p.grad for p in model.parameters()]

    def restore_grads(self) -> None:
        """Re-attach the graph's static gradient tensors (after something replaced ``p.grad``, e.g. an eager step or
        ``zero_grad(set_to_none=True)``)."""
        for p, g in zip(self.model.parameters(), self.static_grads):
            p.grad = g

    def __call__(self, x: Tensor, t: Tensor) -> Tensor:
        if x.shape != self.x.shape or t.shape != self.t.shape:
            raise CartsegError(f"graph captured for {tuple(self.x.shape)} / {tuple(self.t.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss
