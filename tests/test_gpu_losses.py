"""GPU parity: fused losses, metrics and thresholding against golden vectors produced by the
reference's own classes (oracle/make_golden.py) and against the CPU oracle at larger sizes.
fp32 arithmetic on both sides: values to 1e-5 relative, gradients to 1e-4 of the gradient scale;
thresholded masks bit-exact."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden, unpack_bits

pytestmark = pytest.mark.gpu


def _cases():
    import cartseg as cs
    return {
        "bce_dice": cs.BCEDiceLoss(0.5, 1.0),
        "bce_dice_w03_s2": cs.BCEDiceLoss(bce_weight=0.3, smooth=2.0),
        "bce_dice_dims123": cs.BCEDiceLoss(0.5, 1.0, dims=(1, 2, 3)),
        "bce_dice_per_sample": cs.BCEDiceLossPerSample(),
        "focal_a025": cs.FocalLoss(alpha=0.25, gamma=2.0, reduction="mean"),
        "focal_sum_g15": cs.FocalLoss(alpha=0.6, gamma=1.5, reduction="sum"),
        "focal_dice": cs.FocalDiceLoss(alpha=0.5, gamma=2.0, smooth=1.0, w_focal=0.7),
        "boundary": cs.SymmetricBoundaryLoss(),
        "boundary_noabs": cs.SymmetricBoundaryLoss(t=0.4, w_gt=0.8, w_pred=0.3, use_abs=False, scale=2.0),
        "composite": cs.CompositeSegLoss(bce_weight=0.5, boundary_weight=0.3),
    }


NAMES = ["bce_dice", "bce_dice_w03_s2", "bce_dice_dims123", "bce_dice_per_sample", "focal_a025", "focal_sum_g15",
         "focal_dice", "boundary", "boundary_noabs", "composite"]


@pytest.fixture(scope="module")
def loss_inputs():
    g = load_golden("losses.npz")
    shape = tuple(int(v) for v in g["shape"])
    logits = torch.from_numpy(g["logits"])
    targets = torch.from_numpy(unpack_bits(g["targets"], shape).astype(np.float32))
    return g, logits, targets


@pytest.mark.parametrize("name", NAMES)
def test_loss_value_and_grad_vs_reference_golden(loss_inputs, name):
    g, logits, targets = loss_inputs
    crit = _cases()[name]
    x = logits.cuda().requires_grad_(True)
    out = crit(x, targets.cuda())
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[name + "_value"], rtol=2e-5, atol=1e-6)
    (out.sum() if out.dim() else out).backward()
    ref = g[name + "_grad"]
    scale = np.abs(ref).max()
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-4 * scale)


def test_grad_output_scaling_is_applied_on_device(loss_inputs):
    """GradScaler multiplies the loss by 2**16 before backward (train_bce_dice.py:334)."""
    import cartseg as cs
    g, logits, targets = loss_inputs
    x = logits.cuda().requires_grad_(True)
    (cs.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)(x, targets.cuda()) * 65536.0).backward()
    ref = g["focal_dice_grad"] * 65536.0
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())


@pytest.mark.parametrize("B,H,W", [(8, 224, 224), (2, 512, 512)])
def test_losses_vs_oracle_full_size(B, H, W):
    import cartseg as cs
    from oracle import unet_oracle as O
    _, targets = O.synth_batch(B, H, W, seed=3)
    gen = torch.Generator().manual_seed(1)
    logits = torch.randn(B, 1, H, W, generator=gen) * 2.5 + 3.0 * (targets - 0.4)
    pairs = [
        (cs.FocalDiceLoss(0.5, 2.0, 1.0, 0.7), lambda x, t: O.focal_dice_loss(x, t, 0.5, 2.0, 1.0, 0.7)),
        (cs.BCEDiceLoss(), lambda x, t: O.bce_dice_loss(x, t)),
        (cs.CompositeSegLoss(0.5, 0.3), lambda x, t: O.composite_seg_loss(x, t, 0.5, 0.3)),
    ]
    for crit, ref_fn in pairs:
        xr = logits.clone().requires_grad_(True)
        lr = ref_fn(xr, targets)
        lr.backward()
        xg = logits.cuda().requires_grad_(True)
        lg = crit(xg, targets.cuda())
        lg.backward()
        assert lg.item() == pytest.approx(lr.item(), rel=2e-5)
        ref = xr.grad.numpy()
        np.testing.assert_allclose(xg.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())


def test_metrics_and_masks_vs_reference_golden(loss_inputs):
    import cartseg as cs
    g, logits, targets = loss_inputs
    lg, tg = logits.cuda(), targets.cuda()
    assert cs.dice_metric(lg, tg) == pytest.approx(float(g["soft_dice"]), rel=1e-5)
    for t in (0.2, 0.5, 0.65, 0.8):
        tag = f"t{int(round(t * 100)):02d}"
        assert cs.iou_metric(lg, tg, t) == pytest.approx(float(g["iou_" + tag]), rel=1e-5)
        assert cs.hard_dice_metric(lg, tg, t) == pytest.approx(float(g["hard_dice_" + tag]), rel=1e-5)
        np.testing.assert_allclose(cs.precision_recall_f1(lg, tg, t), g["prf_" + tag], rtol=1e-5)
        np.testing.assert_allclose(cs.dice_iou_at_t(lg, tg, t), g["dice_iou_at_" + tag], rtol=1e-5)
        gt = unpack_bits(g["mask_gt_" + tag], logits.shape)
        ge = unpack_bits(g["mask_ge_" + tag], logits.shape)
        from cartseg import ops
        m_gt = torch.ops.cartseg.threshold_mask(lg, ops.logit_bound(t, ge=False)).cpu().numpy().astype(bool)
        m_ge = torch.ops.cartseg.threshold_mask(lg, ops.logit_bound(t, ge=True)).cpu().numpy().astype(bool)
        assert np.array_equal(m_gt, gt), t          # bit-exact thresholded masks
        assert np.array_equal(m_ge, ge), t
        if t == 0.5:
            assert np.array_equal(cs.pseudo_label_mask(lg, t).cpu().numpy().astype(bool), ge[:, 0])
    sw = cs.sweep_thresholds(lg, tg, np.linspace(0.2, 0.8, 13)).cpu().numpy()
    np.testing.assert_allclose(sw, g["sweep13"], rtol=1e-5)


def test_threshold_counts_are_exact_integers_at_full_size():
    import cartseg as cs
    from oracle import unet_oracle as O
    B, H, W = 16, 224, 224
    _, targets = O.synth_batch(B, H, W, seed=9)
    gen = torch.Generator().manual_seed(2)
    logits = torch.randn(B, 1, H, W, generator=gen) * 3 + 2.0 * (targets - 0.5)
    ths = [float(t) for t in torch.linspace(0.05, 0.95, 19)]
    ps, inter, ts, _, _ = cs.threshold_sums(logits.cuda(), targets.cuda(), ths)
    for k, t in enumerate(ths):
        rp, rt, ri = O.hard_counts(logits, targets, t)
        assert torch.equal(ps[:, k].cpu().long(), rp.long())
        assert torch.equal(inter[:, k].cpu().long(), ri.long())
    assert torch.equal(ts.cpu().long(), targets.sum((1, 2, 3)).long())


def test_find_best_threshold_matches_reference_loop():
    import cartseg as cs
    from oracle import unet_oracle as O

    class Fixed(torch.nn.Module):           # stands in for the model: returns pre-baked logits
        def forward(self, x):
            return x

    batches = []
    for s in range(3):
        _, t = O.synth_batch(4, 64, 64, seed=20 + s)
        gen = torch.Generator().manual_seed(s)
        batches.append((torch.randn(4, 1, 64, 64, generator=gen) * 2 + 3 * (t - 0.45), t))
    ths = np.linspace(0.2, 0.8, 13)
    best_t, best_d = cs.find_best_threshold(Fixed(), batches, "cuda", ths)
    ref = np.mean([O.sweep_dice(l, t, ths) for l, t in batches], axis=0)
    assert best_t == pytest.approx(float(ths[int(np.argmax(ref))]))
    assert best_d == pytest.approx(float(ref.max()), rel=1e-5)


def test_losses_reject_cpu_tensors():
    import cartseg as cs
    with pytest.raises(Exception):
        cs.BCEDiceLoss()(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))


def test_focal_loss_reduction_none_vs_oracle():
    """FocalLoss(reduction="none") (src/train_with_focalDice.py:214-219): unreduced map and its gradient against an
    arbitrary element-wise upstream gradient, vs the oracle (fp32 both sides: 2e-5 / 1e-4 of scale)."""
    import cartseg
    from oracle import unet_oracle as O
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(3, 1, 40, 36, generator=g) * 3).requires_grad_(True)
    t = (torch.rand(3, 1, 40, 36, generator=g) > 0.6).float()
    up = torch.randn(3, 1, 40, 36, generator=g)
    for alpha, gamma in ((0.25, 2.0), (0.6, 1.5), (1.0, 0.0)):
        ref = O.focal_loss(x, t, alpha, gamma, reduction="none")
        x.grad = None
        (ref * up).sum().backward()
        xg = x.detach().cuda().requires_grad_(True)
        out = cartseg.FocalLoss(alpha=alpha, gamma=gamma, reduction="none")(xg, t.cuda())
        assert out.shape == ref.shape
        (out * up.cuda()).sum().backward()
        assert (out.detach().cpu() - ref.detach()).abs().max() <= 2e-5 * ref.detach().abs().max() + 1e-7
        assert (xg.grad.cpu() - x.grad).abs().max() <= 1e-4 * x.grad.abs().max()
