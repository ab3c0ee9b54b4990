"""GPU parity of the Active Boundary Loss (cs_abl_forward / cs_abl_backward behind cartseg.ABL) against
(1) golden vectors produced by the reference's own ABL class (oracle/make_golden.py, tests/golden/abl.npz) and
(2) the CPU oracle at training sizes.
Discrete intermediate results are compared exactly (GT distance maps bit-exact; the threshold index and the
predicted-boundary pixel count), the loss to 2e-5 relative and the gradient to 1e-4 of its scale (fp32 both sides)."""
import numpy as np
import pytest
import torch

from test_oracle_abl import abl_golden_cases
from gpu_util import GOLDEN

pytestmark = pytest.mark.gpu


def _run(logits, targets, **kw):
    import cartseg as cs
    from cartseg import ops
    crit = cs.ABL(**kw)
    x = logits.cuda().requires_grad_(True)
    t = targets.cuda()
    loss, valid, scratch = torch.ops.cartseg.abl_loss(x, t.contiguous(), float(crit.label_smoothing),
                                                      float(crit.max_N_ratio), float(crit.max_clip_dist),
                                                      int(crit.ignore_label), bool(crit.per_image_maps))
    dbg = ops.abl_debug(x, scratch, crit.per_image_maps)
    return x, loss, valid, dbg


def test_abl_vs_reference_golden():
    n_cases = 0
    for name, c in abl_golden_cases(GOLDEN):
        x, loss, valid, (eps, k, kept, nb, dmap, kl) = _run(c["logits"], c["targets"])
        B = x.shape[0]
        assert nb == int(c["pred_boundary"].sum()), name              # dilated predicted boundary, pixel count
        assert bool(valid.item()) == (not c["none"]), name
        if c["none"]:
            continue
        assert np.array_equal(dmap.astype(np.float32), c["dist_maps"][:B]), name        # bit-exact distance maps
        np.testing.assert_allclose(float(loss.detach()), float(c["value"]), rtol=2e-5, atol=1e-7, err_msg=name)
        loss.backward()
        scale = max(float(np.abs(c["grad"]).max()), 1e-12)
        np.testing.assert_allclose(x.grad.cpu().numpy(), c["grad"], rtol=1e-3, atol=1e-4 * scale, err_msg=name)
        n_cases += 1
    assert n_cases >= 6


def test_abl_module_returns_none_like_the_reference():
    import cartseg as cs
    t = torch.zeros(2, 1, 32, 32)
    t[:, :, 8:20, 8:20] = 1
    assert cs.ABL()(torch.full_like(t, 0.3).cuda(), t.cuda()) is None           # abl.py:197-198
    crit = cs.BCEDiceABL()
    x = torch.full_like(t, 0.3).cuda().requires_grad_(True)
    total = crit(x, t.cuda())
    region = cs.BCEDiceLoss()(x.detach(), t.cuda())
    assert float(total) == pytest.approx(float(region), rel=1e-6)               # region term alone
    total.backward()
    assert torch.isfinite(x.grad).all()
    assert crit.boundary_none_count == 1 and crit.total_calls == 1


# (1, 1040, 32): taller than the segmented column pass (serial fallback); (2, 16, 1600): wider than the padded row pass
@pytest.mark.parametrize("B,H,W,per_image", [(8, 224, 224, False), (5, 96, 160, False), (4, 128, 128, True),
                                             (1, 1040, 32, False), (2, 16, 1600, False), (3, 50, 70, True)])
def test_abl_vs_oracle_at_training_sizes(B, H, W, per_image):
    from oracle import abl_oracle as A
    from oracle import unet_oracle as O
    _, t = O.synth_batch(B, H, W, seed=31)
    g = torch.Generator().manual_seed(5)
    z = 6.0 * (torch.roll(t, shifts=(3, -4), dims=(2, 3)) - 0.5) + 0.4 * torch.randn(t.shape, generator=g)
    if B > 2:
        t[2] = 0.0                                                   # one image without any GT boundary
    x, loss, valid, (eps, k, kept, nb, dmap, kl) = _run(z, t, per_image_maps=per_image)
    xo = z.clone().requires_grad_(True)
    lo, parts = A.abl_loss(xo, t, per_image_maps=per_image, return_parts=True)
    assert bool(valid.item()) and lo is not None
    assert np.array_equal(dmap.astype(np.float32), parts["dmap"].numpy())
    assert k == parts["k"]
    # a pixel whose KL lies within float rounding of the threshold may flip: allow a handful
    assert abs(nb - int(parts["pred_boundary"].sum())) <= 9 * 3
    assert abs(kept - parts["n_keep"]) <= 9 * 3
    np.testing.assert_allclose(float(loss.detach()), float(lo.detach()), rtol=1e-3)
    if nb == int(parts["pred_boundary"].sum()) and kept == parts["n_keep"]:
        np.testing.assert_allclose(float(loss.detach()), float(lo.detach()), rtol=2e-5)
        lo.backward()
        loss.backward()
        scale = float(xo.grad.abs().max())
        np.testing.assert_allclose(x.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-3, atol=1e-4 * scale)


def test_bce_dice_abl_matches_oracle_and_scales_with_grad_output():
    import cartseg as cs
    from oracle import abl_oracle as A
    from oracle import unet_oracle as O
    _, t = O.synth_batch(4, 64, 64, seed=8)
    g = torch.Generator().manual_seed(9)
    z = 5.0 * (torch.roll(t, shifts=(2, 2), dims=(2, 3)) - 0.5) + 0.3 * torch.randn(t.shape, generator=g)
    x = z.cuda().requires_grad_(True)
    out = cs.BCEDiceABL(bce_weight=0.5, smooth=1.0, abl_weight=0.1)(x, t.cuda())
    xo = z.clone().requires_grad_(True)
    ref = A.bce_dice_abl(xo, t, 0.5, 1.0, 0.1)
    np.testing.assert_allclose(float(out.detach()), float(ref.detach()), rtol=2e-5)
    (out * 1024.0).backward()                                        # GradScaler-style scaling (device scalar)
    ref.backward()
    scale = float(xo.grad.abs().max())
    np.testing.assert_allclose(x.grad.cpu().numpy() / 1024.0, xo.grad.numpy(), rtol=1e-3, atol=1e-4 * scale)
