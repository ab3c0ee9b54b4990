#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code on seeded inputs.

Run in the build container only (needs /root/reference, read-only).  The reference scripts
cannot be imported (they create private directories and import missing packages at import
time), so the hot-path classes/functions are lifted out of their source files by ``ast`` and
``exec``-ed unchanged in a namespace that supplies torch / numpy / scipy.  Nothing from the
reference is written into this repository except the numerical outputs.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O          # noqa: E402  (only for synth inputs / weights)

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def lift(relpath: str, names, extra=None):
    from oracle import ref_lift
    return ref_lift.lift(relpath, names, extra)


def lift_with(relpath, names, extra):
    return lift(relpath, names, extra)


def edge_masks(H, W, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:H, 0:W]
    ms = {}
    ms["all0"] = np.zeros((H, W), bool)
    ms["all1"] = np.ones((H, W), bool)
    m = np.zeros((H, W), bool); m[H // 3, W // 2] = True
    ms["single_fg"] = m
    ms["single_bg"] = ~m
    m = np.zeros((H, W), bool); m[H // 4, :] = True; m[(3 * H) // 4, :] = True   # abl.py:232-234 style
    ms["hlines"] = m
    m = np.zeros((H, W), bool); m[:, W // 5] = True
    ms["vline"] = m
    ms["checker"] = ((yy + xx) % 2).astype(bool)
    ms["bernoulli"] = rng.random((H, W)) < 0.5
    ms["sparse"] = rng.random((H, W)) < 0.01
    ms["disc"] = (yy - H * 0.55) ** 2 + (xx - W * 0.4) ** 2 <= (min(H, W) / 3.1) ** 2
    m = np.zeros((H, W), bool); m[0, 0] = True; m[H - 1, W - 1] = True
    ms["corners"] = m
    m = np.zeros((H, W), bool); m[2:H - 2, 2:W - 2] = True; m[H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = False
    ms["box_with_hole"] = m
    return ms


def grad_probe(n: int, salt: int) -> torch.Tensor:
    """Deterministic closed-form probe vector: a gradient tensor is pinned as a whole by its dot product with it."""
    i = torch.arange(n, dtype=torch.float64)
    return torch.sin(i * 0.7390851332 + 0.113 * salt) + 0.5 * torch.cos(i * 1.1447298858 + 0.071 * salt)


def make_model():
    """The reference's UNet (src/create_testset.py:40-83) + BCEDiceLoss (train_bce_dice.py:186-199) on seeded inputs:
    eval / train logits, loss, and for EVERY parameter the gradient norm, its first 8 elements, its projection on a
    closed-form probe vector (pins the whole tensor) and — for tensors of <= 4096 elements — the full gradient; all BN
    running statistics in full."""
    tb = lift("train_bce_dice.py", ["BCEDiceLoss", "dice_metric", "iou_metric"])
    ct = lift("src/create_testset.py", ["DoubleConv", "UNet"])
    model = {}
    for tag, (B, H, W) in {"a": (2, 32, 32), "b": (1, 48, 16), "c": (3, 64, 96)}.items():
        net = ct["UNet"](in_channels=3, out_channels=1)
        sd = O.synth_state_dict(seed=1)
        net.load_state_dict(sd, strict=True)
        keys = list(net.state_dict().keys())
        assert keys == [k for k, _ in O.state_dict_spec()], "state-dict key order drifted"
        x, tgt = O.synth_batch(B, H, W, seed=5)

        def logits_of(n, inp):                        # forward minus the trailing sigmoid (:83)
            acts = {}
            hnd = n.final_conv.register_forward_hook(lambda m, i, o: acts.__setitem__("z", o))
            n(inp)
            hnd.remove()
            return acts["z"]

        net.eval()
        with torch.no_grad():
            model[f"{tag}_eval_logits"] = logits_of(net, x).numpy()
        net.train()
        z = logits_of(net, x)
        model[f"{tag}_train_logits"] = z.detach().numpy()
        loss = tb["BCEDiceLoss"]()(z, tgt)
        loss.backward()
        model[f"{tag}_train_loss"] = np.float64(loss.item())
        for j, (k, p) in enumerate(net.named_parameters()):
            g = p.grad.detach().double()
            model[f"{tag}_gnorm/{k}"] = np.float64(g.norm().item())
            model[f"{tag}_ghead/{k}"] = p.grad.detach().flatten()[:8].numpy()
            model[f"{tag}_gproj/{k}"] = np.float64(torch.dot(g.flatten(), grad_probe(g.numel(), j)).item())
            if g.numel() <= 4096:
                model[f"{tag}_gfull/{k}"] = p.grad.detach().numpy()
        for k, b in net.named_buffers():
            if k.endswith("running_mean") or k.endswith("running_var"):
                model[f"{tag}_buf/{k}"] = b.detach().flatten()[:8].numpy()
                model[f"{tag}_buffull/{k}"] = b.detach().numpy()
        model[f"{tag}_shape"] = np.array([B, 3, H, W])
    model["n_params"] = np.int64(sum(p.numel() for p in net.parameters()))
    np.savez_compressed(os.path.join(OUT, "model.npz"), **model)


def load_reference_abl():
    """Import the reference's ABL module (src/training/losses/{abl,label_smooth}.py) unchanged, on the CPU.
    Three shims make that possible without touching its arithmetic: ``np.bool`` (removed from numpy, used at
    abl.py:20) is aliased to ``np.bool_``; ``Tensor.cuda()`` (abl.py:87,134-135,193) returns the tensor itself;
    the two files are loaded as a synthetic package so that the relative import at abl.py:8 resolves."""
    import importlib.util
    import types
    import scipy.ndimage  # noqa: F401  (must be imported before np.bool is aliased)
    import torchvision  # noqa: F401
    if not hasattr(np, "bool"):
        np.bool = np.bool_
    torch.Tensor.cuda = lambda self, *a, **k: self
    base = os.path.join(REF, "src", "training", "losses")
    pkg = types.ModuleType("ref_losses")
    pkg.__path__ = [base]
    sys.modules["ref_losses"] = pkg
    mods = {}
    for name in ("label_smooth", "abl"):
        spec = importlib.util.spec_from_file_location("ref_losses." + name, os.path.join(base, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ref_losses." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["abl"]


def abl_cases():
    """name -> (logits [B,1,H,W] fp32, targets [B,1,H,W] fp32 {0,1})"""
    cases = {}
    rng = np.random.Generator(np.random.PCG64(23))

    def noisy(tgt, scale=2.0, gain=4.0):
        z = rng.standard_normal(tgt.shape).astype(np.float32) * scale
        return torch.from_numpy(z) + gain * (tgt - 0.5)

    def shifted(tgt, dy, dx, scale=0.3, gain=6.0):
        """logits of a prediction whose boundary lies (dy, dx) pixels away from the GT boundary"""
        return noisy(torch.roll(tgt, shifts=(dy, dx), dims=(2, 3)), scale, gain)

    _, t = O.synth_batch(1, 48, 64, seed=11)
    cases["b1_48x64"] = (shifted(t, 3, -2), t)
    _, t = O.synth_batch(2, 32, 48, seed=12)
    cases["b2_32x48"] = (shifted(t, -2, 4), t)
    _, t = O.synth_batch(3, 40, 40, seed=13)
    cases["b3_40x40"] = (shifted(t, 5, 5, 0.1, 8.0), t)
    _, t = O.synth_batch(4, 64, 64, seed=14)
    t[1] = 0.0                                             # an image without any GT boundary (empty mask)
    cases["b4_64x64_empty_gt"] = (shifted(t, 4, 0, 0.5, 5.0), t)
    _, t = O.synth_batch(2, 64, 64, seed=16)
    cases["b2_64x64_noisy"] = (noisy(t, 2.0, 4.0), t)     # heavy noise: the threshold ladder climbs far
    _, d = O.synth_batch(1, 64, 64, seed=17)               # no GT boundary at all, B = 1 (scipy's no-zero EDT)
    cases["b1_64x64_no_boundary"] = (noisy(d, 0.2, 6.0), torch.zeros_like(d))
    t = torch.zeros(1, 1, 100, 100); t[0, 0, 5] = 1; t[0, 0, 50] = 1      # abl.py:232-236 smoke shape
    cases["b1_100x100_lines"] = (noisy(t, 1.0, 0.0), t)
    _, t = O.synth_batch(2, 32, 32, seed=15)
    cases["b2_32x32_flat_logits"] = (torch.full_like(t, 0.3), t)          # empty predicted boundary -> None
    return cases


def make_abl():
    import contextlib
    import io
    ref = load_reference_abl()
    out = {}
    for name, (logits, tgt) in abl_cases().items():
        x = logits.clone().requires_grad_(True)
        crit = ref.ABL()
        with contextlib.redirect_stdout(io.StringIO()):
            loss = crit(x, tgt)
            p = torch.sigmoid(x.detach())
            pb = crit.logits2boundary(torch.cat([1 - p, p], 1))
            gb = crit.gt2boundary(tgt[:, 0].long(), ignore_label=crit.ignore_label)
            dm = crit.get_dist_maps(gb)
        out[name + "_logits"] = logits.numpy()
        out[name + "_targets"] = np.packbits(tgt.numpy().astype(bool))
        out[name + "_shape"] = np.array(logits.shape)
        out[name + "_pred_boundary"] = np.packbits(pb.numpy())
        out[name + "_gt_boundary"] = np.packbits(gb.numpy())
        out[name + "_dist_maps"] = dm.numpy().astype(np.int16)            # [2B,H,W], integral values
        if loss is None:
            out[name + "_none"] = np.array(1)
            continue
        out[name + "_none"] = np.array(0)
        out[name + "_value"] = np.float64(loss.item())
        if torch.isfinite(loss):
            loss.backward()
            out[name + "_grad"] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "abl.npz"), **out)


def postproc_masks():
    """name -> uint8 {0,255} masks exercising holes, ties, a foreground corner, enclosed background, emptiness"""
    rng = np.random.Generator(np.random.PCG64(77))
    out = {}
    m = np.zeros((32, 48), np.uint8); m[4:20, 6:30] = 255; m[8:12, 10:16] = 0; m[14, 20] = 0; m[25:29, 38:44] = 255
    out["holes_and_blob"] = m
    m = np.zeros((16, 24), np.uint8); m[1, 0] = m[2, 0] = 255; m[0, 5] = m[0, 6] = 255; m[9:11, 9] = 255
    out["three_way_tie"] = m
    m = np.zeros((16, 24), np.uint8); m[1, 4] = m[2, 4] = 255; m[0, 9] = m[0, 10] = 255
    out["tie_block_order"] = m
    m = np.zeros((24, 24), np.uint8); m[0:6, 0:6] = 255; m[12:20, 12:20] = 255
    out["fg_corner"] = m
    m = np.zeros((24, 32), np.uint8); m[:, 14:17] = 255; m[5:9, 22:28] = 255; m[6:8, 24:26] = 0
    out["stripe_cuts_background"] = m
    out["empty"] = np.zeros((16, 16), np.uint8)
    out["full"] = np.full((16, 16), 255, np.uint8)
    m = np.zeros((20, 20), np.uint8); m[3:9, 3:9] = 255; m[9, 9] = 255; m[10:15, 10:15] = 255; m[11:14, 11:14] = 0
    out["diagonal_touch"] = m                  # 8-connected through one corner; hole closed only 4-connectedly
    m = np.zeros((20, 20), np.uint8); m[2:12, 2:12] = 255; m[4:10, 4:10] = 0; m[10, 10] = 0; m[11, 11] = 0
    out["hole_leaks_diagonally"] = m           # background escapes only through a diagonal: stays a hole (4-conn fill)
    for i, p in enumerate((0.35, 0.5, 0.62)):
        out[f"bernoulli_{i}"] = ((rng.random((64, 64)) < p) * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:224, 0:224]
    blobs = np.zeros((224, 224), bool)
    for _ in range(9):
        cy, cx, r = rng.integers(10, 214), rng.integers(10, 214), rng.integers(4, 40)
        blobs |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
    blobs &= rng.random((224, 224)) > 0.03
    out["blobs_224"] = (blobs * 255).astype(np.uint8)
    m = out["blobs_224"].copy(); m[m == 255] = 200; m[5:9, 5:9] = 100      # grey levels: > 127 vs > 0 binarisation
    out["grey_levels_224"] = m
    return out


def make_postproc():
    import cv2
    cm = lift_with("src/data_preprocessing/clean_masks.py", ["clean_mask"], {"cv2": cv2})
    rb = lift_with("src/data_preprocessing/remove_blops.py", ["clean_mask_largest_component"], {"cv2": cv2})
    pl = lift_with("src/data_preprocessing/create_pseudo_labels_gpu.py", ["entropy_map"], {})
    out = {}
    for name, m in postproc_masks().items():
        out[name + "_mask"] = m
        out[name + "_clean"] = cm["clean_mask"](m.copy())
        out[name + "_largest"] = rb["clean_mask_largest_component"](m.copy())
    # QC scores: the expressions of create_pseudo_labels_gpu.py:294-299 evaluated with the reference's entropy_map
    rng = np.random.Generator(np.random.PCG64(78))
    for name, (H, W) in {"qc_a": (64, 64), "qc_b": (37, 51), "qc_c": (96, 128)}.items():
        _, t = O.synth_batch(1, H if H % 2 == 0 else H + 1, W if W % 2 == 0 else W + 1, seed=H)
        t = t[0, 0, :H, :W].numpy()
        z1 = (rng.standard_normal((H, W)) * 1.5 + 5.0 * (t - 0.5)).astype(np.float32)
        z2 = (rng.standard_normal((H, W)) * 2.5 + 3.0 * (t - 0.4)).astype(np.float32)
        z1[0, :4] = [0.0, 40.0, -40.0, 1e-7]
        w = np.array([0.7, 0.3], dtype=np.float32); w = (w / w.sum()).tolist()             # :167-171
        p = [torch.sigmoid(torch.from_numpy(z)) for z in (z1, z2)]
        probs = (p[0].mul_(w[0])).add_(p[1], alpha=w[1]).numpy()                          # :211-214 (fp32 on the CPU)
        pred01 = (probs >= 0.5).astype(np.uint8)                                          # :294
        out[name + "_logits"] = np.stack([z1, z2])
        out[name + "_probs"] = probs
        out[name + "_pred01"] = np.packbits(pred01)
        out[name + "_fg_area"] = np.float64(float(pred01.mean()))                         # :298
        out[name + "_fg_conf"] = np.float64(float(np.median(np.abs(probs - 0.5) * 2.0)))  # :299
        out[name + "_mean_ent"] = np.float64(float(pl["entropy_map"](probs).mean()))      # :300
    np.savez_compressed(os.path.join(OUT, "postproc.npz"), **out)


def make_preproc():
    """Run the reference's letterbox function and OpenCV's resize on seeded uint8 images; store inputs compactly
    (small images) and the resized uint8 outputs."""
    import cv2
    tb = lift_with("train_bce_dice.py", ["letterbox_image_with_side_padding"], {"SIDE_PADDING_RATIO": 0.1})
    rng = np.random.Generator(np.random.PCG64(91))
    out = {}
    shapes = {"wide": (60, 96), "tall": (120, 50), "square": (64, 64), "tiny_up": (20, 30), "odd": (75, 101),
              "x2": (112, 93)}
    sizes = {"tiny_up": (56, 224), "odd": (56, 224)}
    for name, (H, W) in shapes.items():
        base = rng.integers(0, 256, (H // 4 + 2, W // 4 + 2, 3)).astype(np.uint8)          # smooth-ish content
        img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC)
        img = np.clip(img.astype(int) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
        if name == "x2":                                   # letterboxed side 112 = 2 * 56: the area fast path
            assert tb["letterbox_image_with_side_padding"](img, (0, 0, 0), 0.1).shape[0] == 112
        lb = tb["letterbox_image_with_side_padding"](img, padding_color=(0, 0, 0), side_padding_ratio=0.1)
        out[name + "_image"] = img
        out[name + "_letterbox_side"] = np.array(lb.shape[0])
        for S in sizes.get(name, (56,)):
            out[f"{name}_resized_{S}"] = cv2.resize(lb, (S, S), interpolation=cv2.INTER_LINEAR)
        m = ((rng.random((H, W)) < 0.4) * 255).astype(np.uint8)
        out[name + "_mask"] = np.packbits(m > 0)
        for S in sizes.get(name, (56,)):
            out[f"{name}_mask_{S}"] = np.packbits(cv2.resize(m, (S, S), interpolation=cv2.INTER_NEAREST) > 0)
    np.savez_compressed(os.path.join(OUT, "preproc.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---------------- EDT / SDF -------------------------------------------------------
    bl = lift("src/train_with_boundary_loss.py",
              ["signed_distance_map_np", "batch_sdf_from_masks", "SymmetricBoundaryLoss",
               "CompositeSegLoss", "BCEDiceLoss"])
    sdf = {}
    for (H, W) in ((32, 48), (17, 5), (64, 64)):
        for name, m in edge_masks(H, W, seed=H * 1000 + W).items():
            t = torch.from_numpy(m.astype(np.float32))[None, None]
            sdf[f"{name}_{H}x{W}_mask"] = np.packbits(m)
            sdf[f"{name}_{H}x{W}_raw"] = bl["signed_distance_map_np"](m.astype(np.uint8))
            sdf[f"{name}_{H}x{W}_norm"] = bl["batch_sdf_from_masks"](t)[0, 0].numpy()
    _, m224 = O.synth_batch(2, 224, 224, seed=7)
    sdf["disc_224x224_mask"] = np.packbits(m224[0, 0].numpy().astype(bool))
    sdf["disc_224x224_norm"] = bl["batch_sdf_from_masks"](m224[:1])[0, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "sdf.npz"), **sdf)

    # ---------------- losses & metrics --------------------------------------------------
    tb = lift("train_bce_dice.py", ["BCEDiceLoss", "dice_metric", "iou_metric"])
    fd = lift("src/train_with_focalDice.py", ["FocalLoss", "FocalDiceLoss", "precision_recall_f1"])
    fp = lift("src/finetune_pseudo.py", ["BCEDiceLoss", "dice_metric", "iou_metric"])
    f224 = lift("src/finetune_for_224.py", ["BCEDiceLossPerSample", "dice_iou_at_t"])

    B, H, W = 3, 40, 56
    rng = np.random.Generator(np.random.PCG64(11))
    _, targets = O.synth_batch(B, H, W, seed=3)
    logits = torch.from_numpy((rng.standard_normal((B, 1, H, W)) * 2.5).astype(np.float32))
    logits = logits + 3.0 * (targets - 0.4)          # correlated with the mask, like a trained net
    logits[0, 0, 0, :8] = torch.tensor([0.0, 1e-8, -1e-8, 6e-8, 1.2e-7, -1.2e-7, 30.0, -30.0])
    losses = {"logits": logits.numpy(), "targets": np.packbits(targets.numpy().astype(bool)),
              "shape": np.array([B, 1, H, W])}

    def run(name, crit):
        x = logits.clone().requires_grad_(True)
        out = crit(x, targets)
        if out.dim():
            losses[name + "_value"] = out.detach().numpy()
            out = out.sum()
        else:
            losses[name + "_value"] = np.float64(out.item())
        out.backward()
        losses[name + "_grad"] = x.grad.numpy()

    run("bce_dice", tb["BCEDiceLoss"](bce_weight=0.5, smooth=1.0))
    run("bce_dice_w03_s2", tb["BCEDiceLoss"](bce_weight=0.3, smooth=2.0))
    run("bce_dice_dims123", fp["BCEDiceLoss"](bce_weight=0.5, smooth=1.0))
    run("bce_dice_per_sample", f224["BCEDiceLossPerSample"]())
    run("focal_a025", fd["FocalLoss"](alpha=0.25, gamma=2.0, reduction="mean"))
    run("focal_sum_g15", fd["FocalLoss"](alpha=0.6, gamma=1.5, reduction="sum"))
    run("focal_dice", fd["FocalDiceLoss"](alpha=0.5, gamma=2.0, smooth=1.0, w_focal=0.7))
    run("boundary", bl["SymmetricBoundaryLoss"]())
    run("boundary_noabs", bl["SymmetricBoundaryLoss"](t=0.4, w_gt=0.8, w_pred=0.3, use_abs=False, scale=2.0))
    run("composite", bl["CompositeSegLoss"](bce_weight=0.5, boundary_weight=0.3))

    losses["soft_dice"] = np.float64(tb["dice_metric"](logits, targets))
    for t in (0.2, 0.5, 0.65, 0.8):
        tag = f"t{int(round(t * 100)):02d}"
        losses["iou_" + tag] = np.float64(tb["iou_metric"](logits, targets, t=t))
        losses["hard_dice_" + tag] = np.float64(fp["dice_metric"](logits, targets, t=t))
        losses["prf_" + tag] = np.array(fd["precision_recall_f1"](logits, targets, t=t))
        losses["dice_iou_at_" + tag] = np.array(f224["dice_iou_at_t"](logits, targets, t=t))
        losses["mask_gt_" + tag] = np.packbits((torch.sigmoid(logits) > t).numpy())
        losses["mask_ge_" + tag] = np.packbits((torch.sigmoid(logits) >= t).numpy())
    ths = np.linspace(0.2, 0.8, 13)
    sw = []
    for t in ths:                                         # train_bce_dice.py:223-227
        preds = (torch.sigmoid(logits) > t).float()
        inter = (preds * targets).sum((2, 3)); denom = preds.sum((2, 3)) + targets.sum((2, 3))
        sw.append(((2 * inter + 1.0) / (denom + 1.0)).mean().item())
    losses["sweep13"] = np.array(sw)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **losses)

    make_model()
    make_abl()
    make_postproc()
    make_preproc()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    if len(sys.argv) > 1:                       # e.g. `make_golden.py abl`: regenerate one file only
        for part in sys.argv[1:]:
            globals()["make_" + part]()
    else:
        main()
