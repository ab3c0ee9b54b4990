#!/usr/bin/env python
"""Bandwidth-bound kernels of the hot path and of the 'next' rows, timed alone with CUDA events:
achieved GB/s = ALGORITHMIC bytes per call / average duration, against the measured HBM peak (MEASURED_PEAKS.json),
with the reference's CPU implementation (scipy / OpenCV / numpy through the oracle) timed on a bounded sample beside it.

    python tools/bench_aux.py [--batch 64] [--size 224] [--iters 20]      # one JSON line per kernel group

Inputs are larger than L2 only for some of these (B=64 x 224^2 x 4 B = 12.8 MB per map fits the 126 MB L2), so a
256 MB buffer is written between iterations to flush it; durations are summed per iteration from events around the
call only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
import cartseg as cs                              # noqa: E402
from bench import synth_batch                    # noqa: E402  (oracle/ is imported below for the CPU baselines only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()
B, S = args.batch, args.size
PX = B * S * S
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=args.iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def cpu_time(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def report(name, ms, bytes_per_call, ref_file, cpu_ms=None, cpu_what=None, **extra):
    line = {"kernel": name, "batch": B, "size": S, "ms": ms, "algorithmic_bytes": bytes_per_call,
            "achieved_GBps": bytes_per_call / ms / 1e6, "hbm_peak_GBps": pk, "frac": bytes_per_call / ms / 1e6 / pk,
            "Mpx_per_s": PX / ms / 1e3, "replaces": ref_file}
    if cpu_ms is not None:
        line["cpu_baseline"] = {"ms_per_image": cpu_ms, "what": cpu_what, "cores": 1,
                                "speedup_per_image": cpu_ms / (ms / B)}
    line.update(extra)
    print(json.dumps(line), flush=True)


x, t = synth_batch(min(B, 16), S, S, seed=0)
rep = (B + t.shape[0] - 1) // t.shape[0]
t = t.repeat(rep, 1, 1, 1)[:B].contiguous()
g = torch.Generator().manual_seed(1)
z = 6.0 * (torch.roll(t, shifts=(3, -4), dims=(2, 3)) - 0.5) + 0.5 * torch.randn(t.shape, generator=g)
zd, td = z.cuda(), t.cuda()

# ---- fused loss (focal-Dice): fwd 8 B/px, bwd 12 B/px ------------------------------------------
crit = cs.FocalDiceLoss(alpha=0.5, gamma=2.0, smooth=1.0, w_focal=0.7)
zg = zd.clone().requires_grad_(True)
ms = timed(lambda: crit(zg, td))
report("loss_forward (focal-Dice)", ms, 8 * PX, "src/train_with_focalDice.py:207-235")
loss = crit(zg, td)
ms = timed(lambda: torch.autograd.grad(loss, zg, retain_graph=True))
report("loss_backward (focal-Dice)", ms, 12 * PX, "src/train_with_focalDice.py:207-235")

# ---- exact EDT / SDF: 8 B/px ----------------------------------------------------------------------
ms = timed(lambda: cs.batch_sdf_from_masks(td))
cpu = None
if not args.no_cpu:
    from oracle.unet_oracle import batch_sdf_from_masks as cpu_sdf
    cpu = 1e3 * cpu_time(lambda: cpu_sdf(t[:4]), 1) / 4
report("sdf (exact EDT, columns + rows)", ms, 8 * PX, "src/train_with_boundary_loss.py:191-217", cpu,
       "scipy distance_transform_edt x2 per mask (the reference's own call), 1 thread")

# ---- ABL: fwd reads logits+targets (8), kl map w+r (8), dist map w+r (4), column scratch w+r (8 on half the images)
abl = cs.ABL()
ms = timed(lambda: abl.forward_with_valid(zg, td))
cpu = None
if not args.no_cpu:
    from oracle import abl_oracle as A
    cpu = 1e3 * cpu_time(lambda: A.abl_loss(z[:4], t[:4]), 1) / 4
report("abl_forward (kl + EDT + CE, 4 kernels)", ms, 24 * PX, "src/training/losses/abl.py:66-212", cpu,
       "oracle port of ABL.forward (torch CPU + integer EDT), per image")
la, _ = abl.forward_with_valid(zg, td)
ms = timed(lambda: torch.autograd.grad(la, zg, retain_graph=True))
report("abl_backward", ms, 14 * PX, "src/training/losses/abl.py:66-212")

# ---- threshold stats (13-threshold sweep from one pass): 8 B/px --------------------------------
ths = np.linspace(0.2, 0.8, 13)
ms = timed(lambda: cs.sweep_thresholds(zd, td, ths))
report("threshold_stats (13 thresholds, one pass)", ms, 8 * PX, "train_bce_dice.py:214-232")

# ---- pseudo-label QC: probs read 3x (L2-resident after the first pass), mask 1 B written; algorithmic 5 B/px
probs = torch.sigmoid(zd)[:, 0].contiguous()
ms = timed(lambda: cs.pseudo_label_qc(probs, 0.5))
cpu = None
if not args.no_cpu:
    from oracle import postproc_oracle as P
    pn = probs[:4].cpu().numpy()
    cpu = 1e3 * cpu_time(lambda: [P.qc_scores(p) for p in pn], 1) / 4
report("pseudo_qc (mask + area + exact median + entropy)", ms, 5 * PX,
       "src/data_preprocessing/create_pseudo_labels_gpu.py:294-300", cpu, "numpy threshold/median/entropy per image")

# ---- mask clean-up: 1 B in, 1 B out (algorithmic); union-find scratch traffic on top --------------
mask = (probs >= 0.5).to(torch.uint8) * 255
ms = timed(lambda: cs.clean_mask(mask))
cpu = None
if not args.no_cpu:
    try:
        import cv2
        mn = mask[:8].cpu().numpy()

        def cv_clean():
            for m in mn:
                _, b = cv2.threshold(m, 127, 255, cv2.THRESH_BINARY)
                f = b.copy()
                cv2.floodFill(f, np.zeros((S + 2, S + 2), np.uint8), (0, 0), 255)
                c = cv2.bitwise_or(b, cv2.bitwise_not(f))
                cv2.connectedComponentsWithStats(c, connectivity=8)
        cpu = 1e3 * cpu_time(cv_clean, 3) / 8
    except ImportError:
        cpu = None
report("mask_cleanup (hole fill + largest component)", ms, 2 * PX, "src/data_preprocessing/clean_masks.py:12-32", cpu,
       "OpenCV floodFill + connectedComponentsWithStats per mask")

# ---- input side: uint8 HWC in (3 B per source px), fp32 NCHW out (12 B per output px) ------------
rng = np.random.Generator(np.random.PCG64(0))
src = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(min(B, 8))]
dev = [torch.from_numpy(src[i % len(src)]).cuda() for i in range(B)]
ms = timed(lambda: cs.letterbox_resize_normalize(dev, S))
cpu = None
if not args.no_cpu:
    try:
        import cv2
        from oracle import preproc_oracle as R

        def cv_pre():
            for im in src[:4]:
                lb = R.letterbox(im)
                r = cv2.resize(lb, (S, S), interpolation=cv2.INTER_LINEAR)
                R.normalize_chw(r, cs.preproc.IMAGENET_MEAN, cs.preproc.IMAGENET_STD)
        cpu = 1e3 * cpu_time(cv_pre, 3) / 4
    except ImportError:
        cpu = None
in_bytes = sum(int(d.numel()) for d in dev)
report("preproc_images (letterbox + resize + normalise, 480x640 sources)", ms, in_bytes + 12 * PX,
       "train_bce_dice.py:42-85,147,171-176", cpu, "numpy letterbox + cv2.resize + float32 normalise per image",
       note="includes the host-side descriptor table build + its 40 B/image H2D copy")
