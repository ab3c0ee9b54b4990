#!/usr/bin/env python
"""Per-parameter gradient parity of one training step at a benchmarked size (default K2: B=64, 3x224x224,
focal-Dice): GPU path vs the fp32 CPU oracle and vs the bf16-emulating CPU oracle, listed in BACKWARD order
(final_conv first, conv1 last) so that the table shows where the end-to-end deviation enters.

    python tools/grad_parity.py [--batch 64] [--size 224] [--loss focal_dice] [--out profiles/r2_grad_parity_per_tensor.json]

North-star bar: 3e-2 relative (bf16 path vs the reference's fp32 path).  Columns per tensor:
  gpu_vs_fp32 / gpu_vs_emu   rel-L2 of the GPU gradient against the two oracles (+ cosines)
  emu_vs_fp32                what bf16 storage of activations ALONE does to the oracle (CPU only, fp32 arithmetic)
The oracle is test infrastructure: this tool (like tests/) is a checker, never a product path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))

import torch  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / max(float(a.norm() * b.norm()), 1e-300))


def compute(args):
    import cartseg
    from cartseg import ops
    from oracle import unet_oracle as O

    B, S = args.batch, args.size
    torch.set_num_threads(os.cpu_count() or 1)
    x, tgt = O.synth_batch(B, S, S, seed=5)
    torch.manual_seed(args.seed)
    model = cartseg.UNet()                                         # the reference's default initialisation
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref_fn, crit = {
        "bce_dice": (lambda z, t: O.bce_dice_loss(z, t), cartseg.BCEDiceLoss()),
        "focal_dice": (lambda z, t: O.focal_dice_loss(z, t, 0.5, 2.0, 1.0, 0.7), cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)),
        "composite": (lambda z, t: O.composite_seg_loss(z, t, 0.5, 0.3), cartseg.CompositeSegLoss(0.5, 0.3)),
    }[args.loss]

    model = model.cuda().train()
    z = model(x.cuda())
    loss = crit(z, tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    g_gpu = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    names = [k for k, _ in model.named_parameters()]
    z_gpu = z.detach().cpu()

    def oracle(emulate):
        t0 = time.perf_counter()
        s = {k: v.clone() for k, v in sd.items()}
        keys = O.param_keys(s)
        for k in keys:
            s[k].requires_grad_(True)
        zz = O.unet_logits(x, s, training=True, emulate_bf16=emulate)
        ll = ref_fn(zz, tgt)
        ll.backward()
        return zz.detach(), float(ll), {k: s[k].grad for k in keys}, time.perf_counter() - t0

    z_ref, loss_ref, g_ref, t_ref = oracle(False)
    z_emu, loss_emu, g_emu, t_emu = oracle(True)

    order = [i for st in ops.stage_params() for i in st]           # backward order (stage 0 = head)
    rows = []
    for i in order:
        k = names[i]
        if k.endswith(".conv.0.bias") or k.endswith(".conv.3.bias"):
            # conv bias under a train-mode BN: exactly zero on the GPU (never added), rounding noise ~1e-9 in the oracle
            rows.append({"tensor": k, "numel": g_ref[k].numel(), "note": "conv bias before train-mode BN: gradient is "
                         "identically zero", "gpu_absmax": float(g_gpu[k].abs().max()), "fp32_absmax": float(g_ref[k].abs().max())})
            continue
        rows.append({
            "tensor": k, "numel": g_ref[k].numel(), "fp32_norm": float(g_ref[k].double().norm()),
            "gpu_vs_fp32": rel_l2(g_gpu[k], g_ref[k]), "gpu_vs_emu": rel_l2(g_gpu[k], g_emu[k]),
            "emu_vs_fp32": rel_l2(g_emu[k], g_ref[k]),
            "cos_gpu_fp32": cosine(g_gpu[k], g_ref[k]), "cos_gpu_emu": cosine(g_gpu[k], g_emu[k]),
            "meets_3e-2_vs_fp32": bool(rel_l2(g_gpu[k], g_ref[k]) < 3e-2),
            "meets_3e-2_vs_emu": bool(rel_l2(g_gpu[k], g_emu[k]) < 3e-2),
        })
    live = [r for r in rows if "gpu_vs_fp32" in r]
    keys = [r["tensor"] for r in live]
    cat = lambda g: torch.cat([g[k].double().flatten() for k in keys])   # noqa: E731
    a, f, e = cat(g_gpu), cat(g_ref), cat(g_emu)
    out = {
        "what": f"one training step, B={B}, 3x{S}x{S}, {args.loss}, default init (seed {args.seed}), synthetic batch seed 5",
        "north_star": "loss 1e-2, gradients 3e-2 relative (bf16 path vs the fp32 reference path)",
        "loss": {"gpu": float(loss), "fp32_oracle": loss_ref, "emu_oracle": loss_emu,
                 "rel_vs_fp32": abs(float(loss) - loss_ref) / abs(loss_ref)},
        "logits_rel_l2": {"gpu_vs_fp32": rel_l2(z_gpu, z_ref), "gpu_vs_emu": rel_l2(z_gpu, z_emu), "emu_vs_fp32": rel_l2(z_emu, z_ref)},
        "whole_gradient": {"gpu_vs_fp32": float((a - f).norm() / f.norm()), "gpu_vs_emu": float((a - e).norm() / e.norm()),
                           "emu_vs_fp32": float((e - f).norm() / f.norm()),
                           "cos_gpu_fp32": float(torch.dot(a, f) / (a.norm() * f.norm())),
                           "cos_gpu_emu": float(torch.dot(a, e) / (a.norm() * e.norm())),
                           "cos_emu_fp32": float(torch.dot(e, f) / (e.norm() * f.norm()))},
        "tensors_meeting_3e-2_vs_fp32": sum(r["meets_3e-2_vs_fp32"] for r in live),
        "tensors_meeting_3e-2_vs_emu": sum(r["meets_3e-2_vs_emu"] for r in live),
        "tensors_compared": len(live),
        "cpu_seconds": {"fp32_oracle": t_ref, "emu_oracle": t_emu, "threads": torch.get_num_threads()},
        "per_tensor_backward_order": rows,
    }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--loss", default="focal_dice", choices=["focal_dice", "bce_dice", "composite"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_grad_parity_per_tensor.json"))
    args = ap.parse_args()
    out = compute(args)
    live = [r for r in out["per_tensor_backward_order"] if "gpu_vs_fp32" in r]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: out[k] for k in ("what", "loss", "logits_rel_l2", "whole_gradient", "tensors_meeting_3e-2_vs_fp32",
                                          "tensors_meeting_3e-2_vs_emu", "tensors_compared", "cpu_seconds")}))
    print(f"{'tensor':34s} {'gpu/fp32':>9s} {'gpu/emu':>9s} {'emu/fp32':>9s} {'cos fp32':>9s}")
    for r in live:
        print(f"{r['tensor']:34s} {r['gpu_vs_fp32']:9.4f} {r['gpu_vs_emu']:9.4f} {r['emu_vs_fp32']:9.4f} {r['cos_gpu_fp32']:9.5f}")


if __name__ == "__main__":
    main()
