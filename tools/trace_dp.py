#!/usr/bin/env python
"""Timeline of one DATA-PARALLEL backward pass: the kernels of the two compute streams (cs_unet_trace) next to the
NCCL all-reduces of the communication stream (CUDA events around every bucket) — the "all-reduce overlapped with
wgrad" evidence SURVEY.md §8d asks for, without nsys (not installed in this image).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/trace_dp.py
        [--batch 64] [--size 224] [--wire bf16|fp32] [--bucket-mb 16]

Rank 0 prints the merged timeline and, per all-reduce, what it overlapped with.  Event pairs serialise nothing, but the
per-launch events of the compute streams cost a few microseconds each: read the overlap structure, not the step time.
"""
import argparse
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg                                   # noqa: E402
from cartseg import ops                          # noqa: E402
from bench import synth_batch                   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--wire", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--bucket-mb", type=float, default=16.0)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = cartseg.UNet().to(dev).train()
    cartseg.parallel.init_data_parallel(model, bucket_mb=args.bucket_mb,
                                        wire_dtype=torch.bfloat16 if args.wire == "bf16" else torch.float32)
    sync = ops._DP_STATES[model._dp_handle]
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    x, t = synth_batch(args.batch, args.size, args.size, seed=rank)
    x, t = x.to(dev), t.to(dev)

    def step(mark=None):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), t)
        if mark is not None:
            mark.record()
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    plan = ops.get_plan(args.batch, 3, args.size, args.size, dev, inference_only=False)
    L = cartseg.lib()
    L.cs_unet_trace(plan.handle, 1)
    sync.trace = []
    mark = torch.cuda.Event(enable_timing=True)
    step(mark)
    torch.cuda.synchronize()
    N = 512
    lab, t0, t1 = (C.c_int * N)(), (C.c_double * N)(), (C.c_double * N)()
    n = L.cs_unet_trace_read(plan.handle, N, lab, t0, t1)
    L.cs_unet_trace(plan.handle, 0)
    comm = [(mark.elapsed_time(e0), mark.elapsed_time(e1), numel) for numel, e0, e1 in sync.trace]
    sync.trace = None
    dist.barrier()
    if rank == 0:
        KIND = {10: "fwd_conv", 11: "fwd_bn", 12: "fwd_up", 1: "bn_reduce", 2: "bn_apply", 3: "dgrad", 4: "wgrad", 5: "im2col", 6: "head_bwd", 7: "up_dgrad", 8: "up_wgrad", 9: "up_bias"}
        side = {4, 5, 6, 8, 9}
        # the compute trace is relative to its own first launch; `mark` was recorded just before it on the same stream
        first_kernel = min(t0[i] for i in range(n))
        rows = [(t0[i] - first_kernel, t1[i] - first_kernel, "side" if (lab[i] // 100) in side else "main",
                 f"{KIND.get(lab[i] // 100, '?')} {lab[i] % 100}") for i in range(n)]
        off = min(c[0] for c in comm) if comm else 0.0
        base = min(off, 0.0)
        rows += [(b, e, "comm", f"all_reduce {numel * (2 if args.wire == 'bf16' else 4) / 2**20:.1f} MB") for b, e, numel in comm]
        rows.sort()
        end = max(r[1] for r in rows)
        print(f"# DP backward timeline, rank 0 of {world}, B={args.batch} {args.size}x{args.size}, wire {args.wire}, "
              f"buckets of {args.bucket_mb} MB: {n} compute launches, {len(comm)} all-reduces, {end - base:.3f} ms")
        print("#   begin     end     dur  stream  what")
        for b, e, s, w in rows:
            print(f"{b:9.3f} {e:7.3f} {e - b:7.3f}  {s:5s}  {w}")
        comp = [(b, e, s) for b, e, s, w in rows if s != "comm"]
        tot = sum(e - b for b, e, _ in comm)
        hidden = 0.0
        for b, e, _ in comm:
            # time of this all-reduce during which at least one compute kernel was running
            pts = sorted([(max(b, cb), min(e, ce)) for cb, ce, _ in comp if min(e, ce) > max(b, cb)])
            cur = b
            for pb, pe in pts:
                if pe > cur:
                    hidden += pe - max(pb, cur)
                    cur = pe
        last_compute = max(e for b, e, s in comp)
        print(f"# all-reduce time {tot:.3f} ms, of which {hidden:.3f} ms ran under compute kernels; "
              f"exposed after the last compute kernel: {max(0.0, max(e for b, e, _ in comm) - last_compute):.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
