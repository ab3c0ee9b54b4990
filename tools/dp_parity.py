#!/usr/bin/env python
"""Data-parallel parity on real GPUs: the gradients a rank holds after a DP backward (bucketed NCCL
all-reduce overlapped with the backward kernels) must equal the average of the per-shard gradients computed WITHOUT
any collective — rank 0 recomputes every shard on its own GPU, one after the other, and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_parity.py

BatchNorm statistics are per shard on both sides (DDP semantics; the reference has no SyncBN), so the comparison is
exact up to the fp32 atomics of the split-K weight-gradient accumulation (~1e-6 relative).  Prints one JSON line.

DP_PARITY_BACKEND=gloo runs the same check with every rank on GPU (LOCAL_RANK mod device_count): on a single-GPU box two
ranks share GPU 0 (NCCL refuses two ranks per device; gloo stages CUDA tensors through the host), so the staged backward,
the stage-ordered flat gradient buffer, the bucket slicing and the side-stream joins are exercised without a second GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg                                   # noqa: E402
from bench import synth_batch                   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("DP_PARITY_BACKEND", "nccl")
    local = local % torch.cuda.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)
    B, S = int(os.environ.get("DP_PARITY_BATCH", "8")), int(os.environ.get("DP_PARITY_SIZE", "96"))
    x, t = synth_batch(B * world, S, S, seed=3)                     # the global batch, identical on every rank
    torch.manual_seed(1)
    model = cartseg.UNet().to(dev).train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cartseg.parallel.init_data_parallel(model, bucket_mb=4.0)         # small buckets: several all-reduces in flight
    crit = cartseg.FocalDiceLoss(alpha=0.5, gamma=2.0, smooth=1.0, w_focal=0.7)
    xs = cartseg.parallel.shard_batch(x, rank, world).to(dev)
    ts = cartseg.parallel.shard_batch(t, rank, world).to(dev)
    loss = crit(model(xs), ts)
    loss.backward()
    torch.cuda.synchronize()
    dp_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    loss_mean = loss.detach().clone()
    dist.all_reduce(loss_mean, op=dist.ReduceOp.SUM)
    loss_mean /= world

    # every rank must hold the same averaged gradients
    worst_spread = 0.0
    for k, g in dp_grads.items():
        ref = g.clone()
        dist.broadcast(ref, src=0)
        worst_spread = max(worst_spread, float((g - ref).abs().max()))

    out = None
    if rank == 0:
        single = cartseg.UNet().to(dev).train()
        acc, losses = None, []
        for r in range(world):
            single.load_state_dict(sd0)                               # same weights AND same BN buffers for every shard
            single.zero_grad(set_to_none=True)
            l = crit(single(cartseg.parallel.shard_batch(x, r, world).to(dev)),
                     cartseg.parallel.shard_batch(t, r, world).to(dev))
            l.backward()
            torch.cuda.synchronize()
            losses.append(float(l))
            g = {k: p.grad.detach().clone() for k, p in single.named_parameters()}
            acc = g if acc is None else {k: acc[k] + g[k] for k in g}
        worst, worst_key = 0.0, ""
        num = den = 0.0
        for k in acc:
            ref = acc[k] / world
            e = float((dp_grads[k] - ref).norm() / max(float(ref.norm()), 1e-30))
            num += float((dp_grads[k].double() - ref.double()).pow(2).sum())
            den += float(ref.double().pow(2).sum())
            if e > worst:
                worst, worst_key = e, k
        out = {"check": "dp_parity", "world": world, "backend": backend, "gpus": torch.cuda.device_count(), "per_gpu_batch": B, "size": S,
               "loss_dp_mean": float(loss_mean), "loss_shard_mean": sum(losses) / world,
               "grad_rel_l2_whole": (num / max(den, 1e-300)) ** 0.5, "grad_rel_l2_worst_tensor": worst,
               "worst_tensor": worst_key, "max_abs_spread_between_ranks": worst_spread}
        # fp32 on the wire: exact up to atomics; bf16 on the wire (CARTSEG_DP_WIRE_DTYPE=bf16): two bf16 roundings
        wire = os.environ.get("CARTSEG_DP_WIRE_DTYPE", "fp32")
        out["wire_dtype"] = wire
        tol = 1e-4 if wire == "fp32" else 8e-3
        out["ok"] = bool(out["grad_rel_l2_whole"] < tol and worst_spread == 0.0 and
                         abs(out["loss_dp_mean"] - out["loss_shard_mean"]) < 1e-5 * abs(out["loss_shard_mean"]) + 1e-7)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
