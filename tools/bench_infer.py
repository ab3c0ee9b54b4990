#!/usr/bin/env python
"""Pseudo-label generation throughput (workload K5, BASELINE.json config 5): eval-mode forward (BatchNorm folded into
the conv epilogues) + on-device `sigmoid(logits) >= 0.5` uint8 mask, 3x224x224 inputs, batch sweep.
Reference path: src/data_preprocessing/create_pseudo_labels_gpu.py:201-215,294.  Prints one JSON line per batch size:

  ms / img_per_s          device-resident, eager (one forward = ~32 launches)
  graph_ms / graph_img_per_s   the same forward + threshold replayed from a CUDA graph (cartseg.GraphedInference, weight
                          packs frozen) — what matters at batch 1..16 where the eager path is launch-bound
  e2e_img_per_s           pinned host images in, uint8 masks out (1 B/px D2H instead of the reference's 4 B/px
                          probabilities), serial copies;  e2e_pipelined_img_per_s: H2D of batch i+1 and D2H of batch i-1
                          on side streams under the forward of batch i (cartseg.parallel.CudaPrefetcher)
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg                                   # noqa: E402
from bench import synth_batch                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--graph-max-batch", type=int, default=64)
args = ap.parse_args()
GFLOP_FWD = {224: 73.756, 512: 385.339}[args.size]

torch.manual_seed(0)
model = cartseg.UNet().cuda().eval()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0


def timed_events(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


for B in [int(b) for b in args.batches.split(",")]:
    x, _ = synth_batch(min(B, 8), args.size, args.size, seed=0)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:B].contiguous().cuda()
    x_h = x.cpu().pin_memory()
    iters = args.iters if B >= 16 else 10 * args.iters
    model.freeze_packed(False)

    def run(src):
        with torch.no_grad():
            return cartseg.pseudo_label_mask(model(src), 0.5)

    for _ in range(3):
        run(x)
    ms, m = timed_events(lambda: run(x), iters)
    rec = {"workload": f"pseudo-label inference 3x{args.size}x{args.size}", "batch": B, "ms": ms,
           "img_per_s": B / (ms / 1e3)}
    tf = rec["img_per_s"] * GFLOP_FWD / 1e3
    rec.update(tflops=tf, frac_of_sustained_bf16_peak=tf / peak)

    # ---- CUDA graph replay (weights static: packs frozen, so the graph holds the forward + threshold only)
    if B <= args.graph_max_batch:
        model.freeze_packed(True)
        g = cartseg.GraphedInference(model, x, threshold=0.5)
        gms, gm = timed_events(lambda: g(x), iters)
        assert torch.equal(gm.reshape(m.shape), m)
        rec.update(graph_ms=gms, graph_img_per_s=B / (gms / 1e3),
                   graph_frac_of_sustained_bf16_peak=B / (gms / 1e3) * GFLOP_FWD / 1e3 / peak)
        del g
        model.freeze_packed(False)

    # ---- end to end, serial copies
    out_h = [torch.empty((B, args.size, args.size), dtype=torch.uint8).pin_memory() for _ in range(2)]
    xd = torch.empty_like(x)
    xd.copy_(x_h, non_blocking=True)                     # one untimed pass: first-use costs of the copy paths
    out_h[0].copy_(run(xd).reshape(out_h[0].shape), non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        xd.copy_(x_h, non_blocking=True)
        out_h[0].copy_(run(xd).reshape(out_h[0].shape), non_blocking=True)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.iters
    rec["e2e_img_per_s"] = B / (e2e_ms / 1e3)

    # ---- end to end, pipelined: the loop a pseudo-label generator would write
    class HostBatches:
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            for _ in range(self.n):
                yield (x_h,)

        def __len__(self):
            return self.n

    d2h = torch.cuda.Stream()

    def pipelined(n):
        for i, (xb,) in enumerate(cartseg.parallel.CudaPrefetcher(HostBatches(n), x.device)):
            mask = run(xb)
            d2h.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(d2h):
                out_h[i & 1].copy_(mask.reshape(out_h[0].shape), non_blocking=True)
            mask.record_stream(d2h)
        torch.cuda.synchronize()

    pipelined(2)
    t0 = time.perf_counter()
    pipelined(args.iters)
    pipe_ms = 1e3 * (time.perf_counter() - t0) / args.iters
    rec["e2e_pipelined_img_per_s"] = B / (pipe_ms / 1e3)
    rec["fg_fraction"] = float(m.float().mean().item())
    print(json.dumps(rec), flush=True)
    del xd, out_h, x, x_h, m
    cartseg.ops.release_plans()
    torch.cuda.empty_cache()
