"""Pins the CPU oracle against vectors produced by the reference's own code
(oracle/make_golden.py ran the lifted reference classes; see that script)."""
import os

import numpy as np
import pytest
import torch

from oracle import edt_oracle as E
from oracle import unet_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)


# ------------------------------------------------------------------ EDT / SDF
def test_sdf_bit_exact_vs_reference(golden_dir):
    g = _load(golden_dir, "sdf.npz")
    names = sorted(k[:-5] for k in g if k.endswith("_mask"))
    assert len(names) >= 30
    for n in names:
        H, W = map(int, n.rsplit("_", 1)[1].split("x"))
        m = _unpack(g[n + "_mask"], (H, W))
        if n + "_raw" in g:
            got = E.sdf_of_mask(m)
            assert got.dtype == np.float32
            assert np.array_equal(got.view(np.uint32), g[n + "_raw"].view(np.uint32)), n
        t = torch.from_numpy(m.astype(np.float32))[None, None]
        got = O.batch_sdf_from_masks(t)[0, 0].numpy()
        assert np.array_equal(got.view(np.uint32), g[n + "_norm"].view(np.uint32)), n


def test_edt_matches_scipy_directly():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for H, W, p in ((23, 31, 0.5), (40, 9, 0.9), (9, 40, 0.05), (64, 64, 0.999)):
        m = rng.random((H, W)) < p
        if m.all():
            m[0, 0] = False
        assert np.array_equal(E.edt(m), ndi.distance_transform_edt(m))


def test_sqrt_f32_equals_f64_path_for_all_reachable_values():
    # float32 sqrt of the integer squared distance == float32(float64 sqrt) for every
    # value reachable up to 512x512 — this is what lets the CUDA kernel use sqrtf.
    v = np.arange(0, 2 * 511 * 511 + 1, dtype=np.int64)
    a = np.sqrt(v.astype(np.float32))
    b = np.sqrt(v.astype(np.float64)).astype(np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ------------------------------------------------------------------ losses
LOSS_CASES = {
    "bce_dice": lambda x, t: O.bce_dice_loss(x, t, 0.5, 1.0),
    "bce_dice_w03_s2": lambda x, t: O.bce_dice_loss(x, t, 0.3, 2.0),
    "bce_dice_dims123": lambda x, t: O.bce_dice_loss(x, t, 0.5, 1.0, dims=(1, 2, 3)),
    "bce_dice_per_sample": lambda x, t: O.bce_dice_loss_per_sample(x, t),
    "focal_a025": lambda x, t: O.focal_loss(x, t, 0.25, 2.0, "mean"),
    "focal_sum_g15": lambda x, t: O.focal_loss(x, t, 0.6, 1.5, "sum"),
    "focal_dice": lambda x, t: O.focal_dice_loss(x, t, 0.5, 2.0, 1.0, 0.7),
    "boundary": lambda x, t: O.symmetric_boundary_loss(x, t),
    "boundary_noabs": lambda x, t: O.symmetric_boundary_loss(x, t, t=0.4, w_gt=0.8, w_pred=0.3,
                                                             use_abs=False, scale=2.0),
    "composite": lambda x, t: O.composite_seg_loss(x, t, 0.5, 0.3),
}


@pytest.fixture(scope="module")
def loss_inputs(golden_dir):
    g = _load(golden_dir, "losses.npz")
    shape = tuple(int(v) for v in g["shape"])
    logits = torch.from_numpy(g["logits"])
    targets = torch.from_numpy(_unpack(g["targets"], shape).astype(np.float32))
    return g, logits, targets


@pytest.mark.parametrize("name", sorted(LOSS_CASES))
def test_loss_value_and_grad(loss_inputs, name):
    g, logits, targets = loss_inputs
    x = logits.clone().requires_grad_(True)
    out = LOSS_CASES[name](x, targets)
    np.testing.assert_allclose(out.detach().numpy(), g[name + "_value"], rtol=1e-6, atol=1e-7)
    (out.sum() if out.dim() else out).backward()
    np.testing.assert_allclose(x.grad.numpy(), g[name + "_grad"], rtol=1e-5, atol=1e-9)


def test_metrics_and_masks(loss_inputs):
    g, logits, targets = loss_inputs
    assert O.soft_dice_metric(logits, targets) == pytest.approx(float(g["soft_dice"]), rel=1e-6)
    for t in (0.2, 0.5, 0.65, 0.8):
        tag = f"t{int(round(t * 100)):02d}"
        assert O.iou_metric(logits, targets, t) == pytest.approx(float(g["iou_" + tag]), rel=1e-6)
        assert O.hard_dice_metric(logits, targets, t) == pytest.approx(float(g["hard_dice_" + tag]), rel=1e-6)
        np.testing.assert_allclose(O.precision_recall_f1(logits, targets, t), g["prf_" + tag], rtol=1e-6)
        d, i = g["dice_iou_at_" + tag]
        assert O.hard_dice_metric(logits, targets, t) == pytest.approx(float(d), rel=1e-6)
        assert O.iou_metric(logits, targets, t) == pytest.approx(float(i), rel=1e-6)
        gt = _unpack(g["mask_gt_" + tag], logits.shape)
        ge = _unpack(g["mask_ge_" + tag], logits.shape)
        assert np.array_equal((torch.sigmoid(logits) > t).numpy(), gt)
        assert np.array_equal((torch.sigmoid(logits) >= t).numpy(), ge)
        if t == 0.5:
            assert np.array_equal(O.pseudo_label_mask(logits, t).numpy().astype(bool), ge[:, 0])
    np.testing.assert_allclose(O.sweep_dice(logits, targets, np.linspace(0.2, 0.8, 13)), g["sweep13"], rtol=1e-6)


# ------------------------------------------------------------------ model
def test_state_dict_spec_is_136_keys_31M_params(golden_dir):
    g = _load(golden_dir, "model.npz")
    spec = O.state_dict_spec()
    assert len(spec) == 136
    sd = O.synth_state_dict(seed=1)
    n = sum(sd[k].numel() for k in O.param_keys(sd))
    assert n == int(g["n_params"]) == 31043521
    assert len(O.param_keys(sd)) == 82


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_model_forward_backward(golden_dir, tag):
    g = _load(golden_dir, "model.npz")
    B, C, H, W = (int(v) for v in g[f"{tag}_shape"])
    x, tgt = O.synth_batch(B, H, W, seed=5)
    sd = O.synth_state_dict(seed=1)
    with torch.no_grad():
        z = O.unet_logits(x, sd, training=False)
    np.testing.assert_allclose(z.numpy(), g[f"{tag}_eval_logits"], rtol=1e-4, atol=1e-5)

    keys = O.param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    z = O.unet_logits(x, sd, training=True)
    np.testing.assert_allclose(z.detach().numpy(), g[f"{tag}_train_logits"], rtol=1e-4, atol=1e-5)
    loss = O.bce_dice_loss(z, tgt)
    assert loss.item() == pytest.approx(float(g[f"{tag}_train_loss"]), rel=1e-5)
    loss.backward()
    for k in keys:
        ref = float(g[f"{tag}_gnorm/{k}"])
        got = sd[k].grad.double().norm().item()
        if k.endswith(".conv.0.bias") or k.endswith(".conv.3.bias"):
            # conv bias feeding a train-mode BN: gradient is mathematically 0, numerically noise
            assert got < 1e-5 and ref < 1e-5
            continue
        assert got == pytest.approx(ref, rel=2e-3, abs=1e-7), k
        np.testing.assert_allclose(sd[k].grad.flatten()[:8].numpy(), g[f"{tag}_ghead/{k}"],
                                   rtol=5e-3, atol=1e-6 + 1e-3 * ref / max(1.0, sd[k].numel() ** 0.5))
        # the whole tensor: projection on the closed-form probe of oracle/make_golden.py (+ full copy of small tensors)
        from oracle.make_golden import grad_probe
        probe = grad_probe(sd[k].numel(), keys.index(k))
        proj = torch.dot(sd[k].grad.double().flatten(), probe).item()
        assert proj == pytest.approx(float(g[f"{tag}_gproj/{k}"]), abs=2e-3 * ref * probe.norm().item() + 1e-9), k
        if f"{tag}_gfull/{k}" in g:
            full = g[f"{tag}_gfull/{k}"]
            np.testing.assert_allclose(sd[k].grad.numpy(), full, rtol=5e-3, atol=2e-3 * np.abs(full).max() + 1e-9)
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(sd[k].flatten()[:8].numpy(), g[f"{tag}_buf/{k}"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(sd[k].numpy(), g[f"{tag}_buffull/{k}"], rtol=1e-4, atol=1e-6)
