"""GPU parity at BASELINE.json's full sizes (K2: B=64 @224^2, K3: B=32 @512^2), where the CPU oracle cannot run the
whole batch in seconds: per-sample comparisons where the computation is per-sample (eval-mode forward, EDT), and
size-independent properties elsewhere (exact linearity of the backward pass in dlogits for power-of-two scalings,
run-to-run reproducibility, Dice of thresholded masks)."""
import numpy as np
import pytest
import torch

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def _torch_init_model(seed):
    import cartseg
    torch.manual_seed(seed)
    return cartseg.UNet()


@pytest.mark.parametrize("B,S,check", [(64, 224, (0, 37, 63)), (8, 512, (5,))])
def test_eval_forward_full_batch_matches_oracle_per_sample(B, S, check):
    """Eval-mode forward is per-sample independent (running statistics): samples of a full-size batch are compared
    with the oracle one at a time; masks of those samples agree to Dice >= 0.999."""
    import cartseg
    from oracle import unet_oracle as O
    x, _ = O.synth_batch(B, S, S, seed=11)
    m = _torch_init_model(0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    with torch.no_grad():
        z = m(x.cuda())
        mask = cartseg.pseudo_label_mask(z, 0.5).cpu()
        z = z.cpu()
    for b in check:
        with torch.no_grad():
            ref = O.unet_logits(x[b:b + 1], {k: v.clone() for k, v in sd.items()}, training=False)
        assert rel_l2(z[b:b + 1], ref) < 1e-2, b
        rm = O.pseudo_label_mask(ref, 0.5)[0].bool()
        gm = mask[b].bool()
        if rm.any():
            dice = 2.0 * (rm & gm).sum().item() / (rm.sum().item() + gm.sum().item())
            assert dice >= 0.999, (b, dice)


@pytest.mark.parametrize("B,S", [(64, 224), (32, 512)])
def test_sdf_full_batch(B, S):
    """EDT of a full K4-sized batch: a few images bit-exact against the oracle, and for every image the properties
    that define a signed distance map (sign == class, zero only for degenerate images, |sdf| >= 1/max(H,W))."""
    import cartseg
    from oracle import unet_oracle as O
    _, masks = O.synth_batch(B, S, S, seed=21)
    masks[1] = 0.0
    masks[2] = 1.0
    got = cartseg.batch_sdf_from_masks(masks.cuda()).cpu()
    for b in (0, 1, 2, B - 1):
        ref = O.batch_sdf_from_masks(masks[b:b + 1])
        assert np.array_equal(got[b:b + 1].numpy().view(np.uint32), ref.numpy().view(np.uint32)), b
    fg = masks > 0.5
    regular = torch.ones(B, dtype=torch.bool)
    regular[1] = regular[2] = False
    g = got[regular]
    f = fg[regular]
    assert (g[f] < 0).all() and (g[~f] > 0).all()
    assert (g.abs() >= 1.0 / S - 1e-9).all()
    assert (got[~regular] == 0).all()


def test_train_step_full_k2_properties():
    """K2 (B=64 @224^2, focal-Dice): the loss equals the oracle's loss on the GPU's own logits; the backward pass is
    exactly linear in dlogits for a power-of-two scaling (bf16 and fp32 roundings commute with it — this is what makes
    GradScaler's 2^16 safe); a repeated step reproduces the activation-gradient chain bit for bit."""
    import cartseg
    from oracle import unet_oracle as O
    B, S = 64, 224
    x, tgt = O.synth_batch(B, S, S, seed=31)
    m = _torch_init_model(1).cuda().train()
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    xg, tg = x.cuda(), tgt.cuda()

    def grads(scale):
        m.zero_grad(set_to_none=True)
        z = m(xg)
        loss = crit(z, tg)
        (loss * scale).backward()
        return z.detach(), loss.item(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    z1, l1, g1 = grads(1.0)
    z2, l2, g2 = grads(1.0)
    z3, l3, g3 = grads(65536.0)
    assert torch.equal(z1, z2) and l1 == l2                              # forward bit-reproducible
    ref_loss = O.focal_dice_loss(z1.cpu(), tgt, 0.5, 2.0, 1.0, 0.7).item()
    assert abs(l1 - ref_loss) / ref_loss < 1e-5
    for k in g1:
        if g1[k].abs().max() == 0:
            continue
        assert rel_l2(g2[k], g1[k]) < 1e-4, k                             # fp32 atomic ordering only
        assert rel_l2(g3[k] / 65536.0, g1[k]) < 1e-4, k                   # exact linearity in the loss scale
        assert torch.isfinite(g3[k]).all(), k


def test_train_step_full_k2_teacher_forced_stages_and_loss():
    """Train-mode parity AT the benchmarked size (K2: B=64 @224^2, focal-Dice) — where the weight-gradient split-K
    (K = 3.2 M pixels) and the fp64 BN-statistics atomics operate.  Loss vs the oracle on the GPU's logits; then the
    full-resolution blocks conv1.* / dconv1.* and the bottleneck conv5.* stage by stage ("teacher-forced": each kernel's
    output against torch-CPU fp32 applied to the tensors the GPU path actually produced, read back through
    cs_unet_debug_read): conv fprop, BN statistics + ReLU (+pool), BN backward, BN parameter gradients, wgrad, dgrad;
    the conv-transposes feeding dconv1 / leaving conv5, and the head.  Bars as in test_gpu_unet_stages.py
    (4e-3 bf16 tensors, 3e-3 fp32 parameter gradients; the north-star bar is 3e-2).  ~1 min of CPU on 16 cores."""
    import cartseg
    from cartseg import ops
    from oracle import unet_oracle as O
    import test_gpu_unet_stages as S
    B, H = 64, 224
    x, tgt = O.synth_batch(B, H, H, seed=5)
    torch.manual_seed(0)
    m = cartseg.UNet()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    z, loss = S._run(m, crit, x, tgt)
    ref_loss = O.focal_dice_loss(z.detach().cpu(), tgt, 0.5, 2.0, 1.0, 0.7).item()
    assert abs(loss.item() - ref_loss) / ref_loss < 1e-5
    plan = ops.get_plan(B, 3, H, H, torch.device("cuda"), inference_only=False)
    snap = S._LazySnapshot(plan)
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    rows = S.check_stages(snap, sd, x, z.detach().cpu(), z.grad.detach().cpu(), grads,
                          conv_ids=(0, 1, 8, 9, 16, 17), up_ids=(0, 3), head=True)
    worst = sorted(rows, key=lambda r: -r[1] / r[2])[:8]
    print(f"\n[K2 B{B} {H}x{H}] {len(rows)} stage checks; worst (rel-L2 / tolerance):")
    for n, e, tol in worst:
        print(f"   {n:40s} {e:.3e} / {tol:g}")
    bad = [(n, f"{e:.3e}", tol) for n, e, tol in rows if not e < tol]
    assert not bad, bad


def test_gradients_full_k2_per_tensor_vs_both_oracles():
    """End-to-end gradients AT K2 size (B=64 @224^2, focal-Dice, default init) against the fp32 oracle and the
    bf16-emulating oracle, per tensor (tools/grad_parity.py; table committed as profiles/r2_grad_parity_per_tensor.json).
    What holds, and is asserted:
      * loss 1e-2 (measured 7e-6); logits 3e-2 (measured 1.1e-2);
      * the ten tensors nearest the loss (final_conv, dconv1.*, upconv1) meet the north-star 3e-2 against the FP32
        oracle (measured 1e-3 ... 2.4e-2);
      * every tensor is as close to the fp32 oracle as the oracle's own bf16-storage emulation is (ratio <= 1.1): the
        remaining deviation (up to 0.46 rel-L2 at the bottleneck, 14 layers from the loss) is what rounding stored
        activations to bf16 does to ANY implementation of this network — a CPU-only fact (emu_vs_fp32) — not a kernel
        error; DESIGN.md §1 states this deviation from north_star's 3e-2;
      * whole gradient: cosine >= 0.99 vs fp32 (measured 0.9964), >= 0.997 vs the emulation (0.9989)."""
    import argparse
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import grad_parity
    out = grad_parity.compute(argparse.Namespace(batch=64, size=224, loss="focal_dice", seed=0))
    print({k: out[k] for k in ("loss", "logits_rel_l2", "whole_gradient", "tensors_meeting_3e-2_vs_fp32")})
    assert out["loss"]["rel_vs_fp32"] < 1e-2
    assert out["logits_rel_l2"]["gpu_vs_fp32"] < 3e-2
    wg = out["whole_gradient"]
    assert wg["cos_gpu_fp32"] >= 0.99 and wg["cos_gpu_emu"] >= 0.997, wg
    assert wg["gpu_vs_fp32"] <= 1.1 * wg["emu_vs_fp32"] + 5e-3, wg
    near = ("final_conv.", "dconv1.", "upconv1.")
    for r in out["per_tensor_backward_order"]:
        if "gpu_vs_fp32" not in r:
            assert r["gpu_absmax"] == 0.0, r                       # conv bias before a train-mode BN
            continue
        if r["tensor"].startswith(near):
            assert r["gpu_vs_fp32"] < 3e-2, r
        if r["tensor"].startswith("upconv") and r["tensor"].endswith(".bias"):
            # Sum over pixels of an activation gradient that cancels to ~1e-3 of its terms (the BN backward above it
            # removes the mean; only image-border pixels contribute): the bf16 storage of that gradient — which the
            # emulation, with fp32 gradients, does not have — shows up here.  Measured 2.4e-2 ... 1.3e-1.
            assert r["gpu_vs_fp32"] < 0.2, r
            continue
        assert r["gpu_vs_fp32"] <= 1.1 * r["emu_vs_fp32"] + 1e-2, r
        assert r["gpu_vs_emu"] <= 0.8 * r["emu_vs_fp32"] + 2e-2, r
