#!/usr/bin/env python
"""One training step of the bench workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_step.py
(also prints the CUDA-event time of the same step when run without ncu)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg                                   # noqa: E402
from bench import synth_batch                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--loss", default="focal_dice")
ap.add_argument("--eval", action="store_true")
args = ap.parse_args()

torch.manual_seed(0)
model = cartseg.UNet().cuda()
crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7) if args.loss == "focal_dice" else cartseg.CompositeSegLoss(0.5, 0.3)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
x, t = synth_batch(args.batch, args.size, args.size, seed=0)
x, t = x.cuda(), t.cuda()


def step():
    if args.eval:
        with torch.no_grad():
            return model(x)
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    opt.step()
    return loss


model.train(not args.eval)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"one step: {e0.elapsed_time(e1):.3f} ms")
