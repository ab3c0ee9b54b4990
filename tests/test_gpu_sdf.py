"""GPU parity: exact EDT / signed distance maps (cartseg::sdf -> cs_sdf) — BIT-EXACT against the
reference's scipy path (golden vectors produced by the reference's own signed_distance_map_np /
batch_sdf_from_masks, src/train_with_boundary_loss.py:191-217) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden, unpack_bits

pytestmark = pytest.mark.gpu


def _gpu_sdf(mask_bool: np.ndarray) -> np.ndarray:
    import cartseg
    t = torch.from_numpy(mask_bool.astype(np.float32)).cuda()
    while t.dim() < 4:
        t = t[None]
    return cartseg.batch_sdf_from_masks(t).cpu().numpy()


def test_golden_masks_bit_exact():
    g = load_golden("sdf.npz")
    names = sorted(k[:-5] for k in g if k.endswith("_mask"))
    assert len(names) >= 30
    bad = []
    for n in names:
        H, W = map(int, n.rsplit("_", 1)[1].split("x"))
        m = unpack_bits(g[n + "_mask"], (H, W))
        got = _gpu_sdf(m)[0, 0]
        if not np.array_equal(got.view(np.uint32), g[n + "_norm"].view(np.uint32)):
            bad.append((n, int((got.view(np.uint32) != g[n + "_norm"].view(np.uint32)).sum())))
    assert not bad, bad


# (1040, 48): taller than the segmented column pass (serial fallback); (8, 2112): wider than the padded row pass;
# (257, 33) / (300, 250): segment counts 16 with ragged last segments and ragged 32-column groups
@pytest.mark.parametrize("H,W,B", [(224, 224, 6), (512, 512, 2), (96, 160, 3), (16, 16, 5), (1040, 48, 2),
                                   (8, 2112, 2), (257, 33, 5), (300, 250, 3), (1024, 64, 2)])
def test_batches_bit_exact_vs_oracle(H, W, B):
    from oracle import unet_oracle as O
    rng = np.random.default_rng(H * 7 + W)
    masks = np.zeros((B, 1, H, W), dtype=np.float32)
    _, disc = O.synth_batch(B, H, W, seed=11)
    masks[:] = disc.numpy()
    if B > 1:
        masks[1, 0] = (rng.random((H, W)) < 0.5)            # Bernoulli noise
    if B > 2:
        masks[2, 0] = 0.0                                   # empty image -> zeros (:200-201)
    if B > 3:
        masks[3, 0] = 1.0                                   # full image  -> zeros
    if B > 4:
        masks[4, 0] = 0.0
        masks[4, 0, H // 2, W // 3] = 1.0                   # single pixel
    ref = O.batch_sdf_from_masks(torch.from_numpy(masks)).numpy()
    got = _gpu_sdf(masks.astype(bool))
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_threshold_semantics_and_pred_sdf():
    """SDF of the thresholded prediction: (sigmoid(x) > t).float() -> batch_sdf (:250-251)."""
    import cartseg
    from cartseg import ops
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    logits = torch.randn(3, 1, 64, 80) * 2
    logits[0, 0, 0, :4] = torch.tensor([0.0, 1e-8, 6e-8, 1.2e-7])
    for t in (0.5, 0.4, 0.73):
        ref = O.batch_sdf_from_masks((torch.sigmoid(logits) > t).float()).numpy()
        got = torch.ops.cartseg.sdf(logits.cuda(), ops.logit_bound(t, ge=False), True, 80.0).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), t


def test_rejects_cpu_tensors():
    import cartseg
    with pytest.raises(Exception):
        cartseg.batch_sdf_from_masks(torch.zeros(1, 1, 16, 16))
