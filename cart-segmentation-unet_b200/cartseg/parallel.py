"""Data-parallel training: one process per GPU, batch sharded across ranks, gradients all-reduced
over NCCL (NVLink 5 / NVSwitch) while the backward pass is still running.

The reference is single-device (no torch.distributed anywhere — SURVEY.md §2.3); this is the one
parallel strategy the path admits: images are independent units, the only exchange step is the
gradient average.  BatchNorm statistics stay per shard (standard DDP semantics).

How the overlap works: ``cs_unet_backward`` runs the 23 backward stages in reverse execution order and
writes parameter gradients into ONE flat buffer laid out in that same order (ops.grad_layout).  Stages
are grouped into buckets of ~``bucket_mb``; as soon as the kernels of a bucket are enqueued, an event is
recorded and the bucket's contiguous slice is all-reduced on a side stream while the compute stream
continues with the next stages.  The compute stream joins the side stream once, after the last stage.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# Bucket size of the gradient all-reduce.  Measured on 2 x B200 in one box (round 1, k2, ms per step): 16 MB 19.01 / 19.04,
# 4 MB 19.18 / 19.20, 2 MB 19.50, one bucket after the whole backward (no overlap) 19.52; single GPU 18.8.
# On 8 x B200 (round 2, k2, one box, tools/dp8_buckets.sh): 16 MB 17.49, 48 MB 17.39, ONE bucket after the backward pass
# 17.26 ms — with eight ranks every overlapped all-reduce holds SMs that the single-wave persistent GEMM grids then
# miss (a kernel launched while NCCL is resident runs a second, nearly empty wave), which costs more than the ~0.8 ms the
# exposed 124 MB all-reduce takes.  Hence: overlapped 16 MB buckets up to 2 ranks, one bucket from 4 ranks on.
DEFAULT_BUCKET_MB = 16.0
LARGE_WORLD_BUCKET_MB = 1.0e5


def default_bucket_mb(world: int) -> float:
    return DEFAULT_BUCKET_MB if world <= 2 else LARGE_WORLD_BUCKET_MB


def plan_buckets(stage_off: Sequence[int], bucket_elems: int) -> List[Tuple[int, int]]:
    """Group consecutive backward stages into buckets of at least ``bucket_elems`` gradient elements.
    ``stage_off[s]`` is the flat offset at which stage ``s`` starts (len = stages + 1).
    Returns [(stage_begin, stage_end), ...] covering every stage exactly once, in order."""
    n = len(stage_off) - 1
    out: List[Tuple[int, int]] = []
    s0 = 0
    for s in range(n):
        if stage_off[s + 1] - stage_off[s0] >= bucket_elems or s == n - 1:
            out.append((s0, s + 1))
            s0 = s + 1
    return out


class GradSync:
    """Averages slices of the flat gradient buffer across the process group, asynchronously."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, bucket_mb: float = 16.0,
                 wire_dtype: torch.dtype = torch.float32):
        """``wire_dtype=torch.bfloat16`` sends the gradients over NVLink as bf16 (62 MB instead of 124 MB per step for
        this network): the NCCL kernels then hold their SMs for half as long next to the single-wave persistent GEMM
        grids.  The fp32 gradients are rounded once before and the bf16 average once after the collective (relative
        error <= 2^-8 per element; the optimizer still sees fp32 tensors)."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if wire_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("wire_dtype must be torch.float32 or torch.bfloat16")
        self.group = group
        self.world = dist.get_world_size(group)
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20) / 4))
        self.wire_dtype = wire_dtype
        self._comm_stream = None
        self._works = []
        self._use_avg = dist.get_backend(group) == "nccl"
        self._pending_scale: List[torch.Tensor] = []
        self.trace = None                      # list of (label, begin event, end event) when tracing is on

    def stage_buckets(self, stage_off: Sequence[int]) -> List[Tuple[int, int]]:
        return plan_buckets(stage_off, self.bucket_elems)

    def reduce_async(self, flat_slice: torch.Tensor, wait=None) -> None:
        """All-reduce (average) ``flat_slice`` on the communication stream once the work that produces it is done.
        ``wait(cuda_stream_handle)``, if given, makes that raw stream wait for the producer's internal streams
        (cs_unet_backward_wait); the stream also waits for everything enqueued on the current stream so far."""
        if flat_slice.numel() == 0 or self.world == 1:
            return
        if flat_slice.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat_slice.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat_slice.device))
            self._comm_stream.wait_event(ev)
            if wait is not None:
                wait(self._comm_stream.cuda_stream)
            with torch.cuda.stream(self._comm_stream):
                if self.trace is not None:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record(self._comm_stream)
                op = dist.ReduceOp.AVG if self._use_avg else dist.ReduceOp.SUM
                if self.wire_dtype is torch.bfloat16:
                    wire = flat_slice.to(torch.bfloat16)
                    wire.record_stream(self._comm_stream)
                    dist.all_reduce(wire, op=op, group=self.group)
                    flat_slice.copy_(wire)
                else:
                    dist.all_reduce(flat_slice, op=op, group=self.group)
                if not self._use_avg:
                    flat_slice.mul_(1.0 / self.world)
                if self.trace is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record(self._comm_stream)
                    self.trace.append((flat_slice.numel(), e0, e1))
        else:                                           # CPU tensors (gloo): used by the host-logic tests
            self._works.append(dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._pending_scale.append(flat_slice)

    def finish(self) -> None:
        """Make the reduced gradients visible to the current stream (call once per backward)."""
        if self._comm_stream is not None:
            torch.cuda.current_stream(self._comm_stream.device).wait_stream(self._comm_stream)
        for w in self._works:
            w.wait()
        for t in self._pending_scale:
            t.mul_(1.0 / self.world)
        self._works.clear()
        self._pending_scale.clear()


_next_handle = 1


def init_data_parallel(model, group: Optional[dist.ProcessGroup] = None, bucket_mb: Optional[float] = None,
                       broadcast_from: int = 0, reserve_sms: int = 0, wire_dtype: Optional[torch.dtype] = None):
    """Make ``model`` (a cartseg.UNet) data-parallel over ``group``: broadcast rank-``broadcast_from``'s
    parameters and BN buffers, and hook the bucketed gradient all-reduce into its backward.  Returns the model.

    ``reserve_sms`` > 0 caps the persistent GEMM grids at ``SMs - reserve_sms`` so that SMs held by NCCL kernels do
    not push GEMM CTAs into a second wave.  Measured on 8 x B200 (round 1) it did not pay: 21.3 ms per step with 4
    reserved SMs and NCCL_MAX_NCHANNELS=4 against 20.7 ms with the defaults, so the default is 0."""
    global _next_handle
    import os
    from . import ops
    if bucket_mb is None:
        bucket_mb = float(os.environ.get("CARTSEG_DP_BUCKET_MB", default_bucket_mb(dist.get_world_size(group))))
    if not reserve_sms:
        reserve_sms = int(os.environ.get("CARTSEG_DP_RESERVE_SMS", "0"))
    if wire_dtype is None:
        wire_dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[os.environ.get("CARTSEG_DP_WIRE_DTYPE", "fp32")]
    sync = GradSync(group, bucket_mb, wire_dtype)
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t, src=broadcast_from, group=group)
    handle = _next_handle
    _next_handle += 1
    ops._DP_STATES[handle] = sync
    model._dp_handle = handle
    model._reserve_sms = max(0, int(reserve_sms)) if dist.get_world_size(group) > 1 else 0
    return model


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous equal shard of a global batch (the batch must divide evenly: equal shards are what
    makes mean-of-shard-losses equal the global-batch loss)."""
    if x.shape[0] % world:
        raise ValueError(f"global batch {x.shape[0]} is not divisible by world size {world}")
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]


class CudaPrefetcher:
    """Wraps an iterable of (pinned) host batches — e.g. ``DataLoader(..., pin_memory=True)`` as every training script
    of the reference builds it (train_bce_dice.py:284-287) — and stages batch i+1 on the device with a side stream
    while step i computes, so the host->device copy of ``data.to(DEVICE)`` (train_bce_dice.py:329) leaves the
    critical path.  Two sets of device buffers are reused in turn (no allocator traffic per step); a yielded batch
    stays valid until the batch after the next one is requested."""

    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = torch.device(device if device is not None else "cuda")
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]
        self._free = [None, None]           # event: the consumer has finished with the slot's previous contents

    def _stage(self, batch, slot):
        bufs = self._slots[slot]
        if bufs is None or len(bufs) != len(batch) or any(
                torch.is_tensor(t) and (b is None or b.shape != t.shape or b.dtype != t.dtype)
                for t, b in zip(batch, bufs)):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) if torch.is_tensor(t) else None for t in batch]
            self._slots[slot] = bufs
        if self._free[slot] is not None:
            self.stream.wait_event(self._free[slot])
        with torch.cuda.stream(self.stream):
            out = tuple(b.copy_(t, non_blocking=True) if torch.is_tensor(t) else t for t, b in zip(batch, bufs))
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        slot = 0
        try:
            nxt = self._stage(next(it), slot)
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            cur_slot = slot
            slot ^= 1
            try:
                nxt = self._stage(next(it), slot)
            except StopIteration:
                nxt = None
            cs = torch.cuda.current_stream(self.device)
            cs.wait_event(ev)
            yield cur
            done = torch.cuda.Event()          # everything the consumer enqueued on its stream for this batch
            done.record(torch.cuda.current_stream(self.device))
            self._free[cur_slot] = done

    def __len__(self):
        return len(self.loader)
