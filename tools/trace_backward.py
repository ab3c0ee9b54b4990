#!/usr/bin/env python
"""Timeline of one backward pass (cs_unet_trace): which kernels of the two internal streams really overlap.
    python tools/trace_backward.py [--batch 64] [--size 224]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg                                   # noqa: E402
from cartseg import ops                          # noqa: E402
from bench import synth_batch                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=224)
args = ap.parse_args()
torch.manual_seed(0)
model = cartseg.UNet().cuda().train()
crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
x, t = synth_batch(args.batch, args.size, args.size, seed=0)
x, t = x.cuda(), t.cuda()


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), t)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
plan = ops.get_plan(args.batch, 3, args.size, args.size, x.device, inference_only=False)
L = cartseg.lib()
L.cs_unet_trace(plan.handle, 1)
step()
torch.cuda.synchronize()
N = 512
lab, t0, t1 = (C.c_int * N)(), (C.c_double * N)(), (C.c_double * N)()
n = L.cs_unet_trace_read(plan.handle, N, lab, t0, t1)
L.cs_unet_trace(plan.handle, 0)
KIND = {10: "fwd_conv", 11: "fwd_bn", 12: "fwd_up", 1: "bn_reduce", 2: "bn_apply", 3: "dgrad", 4: "wgrad", 5: "im2col", 6: "head_bwd", 7: "up_dgrad", 8: "up_wgrad", 9: "up_bias"}
fwd = sorted((t0[i], t1[i], lab[i]) for i in range(n) if lab[i] >= 1000)
if fwd:
    print(f"# forward timeline (one stream), {len(fwd)} launches, {fwd[-1][1] - fwd[0][0]:.3f} ms from first launch to last completion")
    print("#   begin     end     dur   gap-before  kernel")
    prev = fwd[0][0]
    for b, e, l in fwd:
        print(f"{b - fwd[0][0]:9.3f} {e - fwd[0][0]:7.3f} {e - b:7.3f}  {b - prev:7.3f}     {KIND.get(l // 100, str(l // 100)):9s} {l % 100:2d}")
        prev = e
    print(f"# forward kernel time {sum(e - b for b, e, l in fwd):.3f} ms, gaps (incl. untraced stem / loss kernels) {fwd[-1][1] - fwd[0][0] - sum(e - b for b, e, l in fwd):.3f} ms")
rows = sorted((t0[i], t1[i], lab[i]) for i in range(n) if lab[i] < 1000)
base = rows[0][0]
rows = [(b - base, e - base, l) for b, e, l in rows]
n = len(rows)
end = max(r[1] for r in rows)
print(f"# backward timeline, {n} launches, {end:.3f} ms from first launch to last completion")
print("#   begin     end     dur  stream  kernel        overlap with the other stream (ms)")
side = {4, 5, 6, 8, 9}
for b, e, l in rows:
    k, idx = l // 100, l % 100
    other = [(b2, e2) for b2, e2, l2 in rows if ((l2 // 100) in side) != (k in side)]
    ov = sum(max(0.0, min(e, e2) - max(b, b2)) for b2, e2 in other)
    print(f"{b:9.3f} {e:7.3f} {e - b:7.3f}  {'side' if k in side else 'main'}    {KIND.get(k, str(k)):10s} {idx:2d}   {ov:6.3f}")
main_busy = sum(e - b for b, e, l in rows if (l // 100) not in side)
side_busy = sum(e - b for b, e, l in rows if (l // 100) in side)
print(f"# main-stream kernel time {main_busy:.3f} ms, side-stream kernel time {side_busy:.3f} ms, wall {end:.3f} ms")
