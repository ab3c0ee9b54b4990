"""GPU parity of the pseudo-label post-processing (cs_ensemble_accumulate, cs_pseudo_qc, cs_mask_cleanup) against
golden vectors produced by the reference's own functions with OpenCV (tests/golden/postproc.npz) and against the CPU
oracle at full size.  Masks, foreground counts and the median confidence are bit-exact; ensemble probabilities to
2 float32 ulps (device expf vs libm); mean entropy to 1e-6 absolute."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden

pytestmark = pytest.mark.gpu


def test_mask_cleanup_bit_exact_vs_reference_golden():
    import cartseg as cs
    g = load_golden("postproc.npz")
    names = sorted(k[:-5] for k in g if k.endswith("_mask"))
    assert len(names) >= 14
    for n in names:
        m = torch.from_numpy(g[n + "_mask"]).cuda()
        got = cs.clean_mask(m).cpu().numpy()
        assert np.array_equal(got, g[n + "_clean"]), n
        got = cs.clean_mask_largest_component(m).cpu().numpy()
        want = g[n + "_largest"]
        assert np.array_equal(got, want if want.max() > 1 else want * 0), n


def test_mask_cleanup_batched_equals_per_image():
    import cartseg as cs
    from oracle import postproc_oracle as P
    rng = np.random.Generator(np.random.PCG64(3))
    yy, xx = np.mgrid[0:224, 0:224]
    masks = []
    for b in range(16):
        m = np.zeros((224, 224), bool)
        for _ in range(int(rng.integers(1, 8))):
            cy, cx, r = rng.integers(0, 224), rng.integers(0, 224), rng.integers(3, 60)
            m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        m &= rng.random((224, 224)) > 0.02 * (b % 4)
        masks.append((m * 255).astype(np.uint8))
    masks[5][:] = 0
    masks[6][0, 0] = 255
    batch = torch.from_numpy(np.stack(masks)).cuda()
    got_clean = cs.clean_mask(batch).cpu().numpy()
    got_big = cs.clean_mask_largest_component(batch).cpu().numpy()
    for b, m in enumerate(masks):
        assert np.array_equal(got_clean[b], P.clean_mask(m)), b
        assert np.array_equal(got_big[b], P.clean_mask_largest_component(m)), b


def test_mask_cleanup_worst_case_shapes():
    """Serpentine and checkerboard masks: long union-find chains, thousands of components."""
    import cartseg as cs
    from oracle import postproc_oracle as P
    H = W = 128
    snake = np.zeros((H, W), np.uint8)
    snake[::2, :] = 255
    snake[1::4, -1] = 255
    snake[3::4, 0] = 255
    checker = (((np.add.outer(np.arange(H), np.arange(W))) % 2) * 255).astype(np.uint8)
    rows = np.zeros((H, W), np.uint8); rows[::2, 1:-1] = 255
    for m in (snake, checker, rows):
        t = torch.from_numpy(m).cuda()
        assert np.array_equal(cs.clean_mask(t).cpu().numpy(), P.clean_mask(m))
        assert np.array_equal(cs.clean_mask_largest_component(t).cpu().numpy(), P.clean_mask_largest_component(m))


def test_qc_scores_vs_reference_golden():
    import cartseg as cs
    g = load_golden("postproc.npz")
    for n in ("qc_a", "qc_b", "qc_c"):
        z = torch.from_numpy(g[n + "_logits"]).cuda()                    # [2,H,W]
        H, W = z.shape[1:]
        models = [lambda x, i=i: z[i][None, None] for i in range(2)]
        w = cs.postproc.normalize_weights([0.7, 0.3])
        probs = cs.ensemble_forward(models, w, torch.zeros(1, 3, H, W, device="cuda"))
        np.testing.assert_allclose(probs[0].cpu().numpy(), g[n + "_probs"], rtol=0, atol=2.4e-7, err_msg=n)
        # scores on the reference's own probabilities: mask, count and median are exact
        p_ref = torch.from_numpy(g[n + "_probs"]).cuda()[None]
        mask, fg_area, fg_conf, mean_ent = cs.pseudo_label_qc(p_ref, 0.5)
        want = np.unpackbits(g[n + "_pred01"])[:H * W].reshape(H, W)
        assert np.array_equal(mask[0].cpu().numpy(), want), n
        assert float(fg_area[0]) == float(g[n + "_fg_area"]), n
        assert float(fg_conf[0]) == float(g[n + "_fg_conf"]), n
        assert abs(float(mean_ent[0]) - float(g[n + "_mean_ent"])) <= 1e-6, n


@pytest.mark.parametrize("B,H,W", [(64, 224, 224), (3, 101, 77), (2, 512, 512)])
def test_qc_scores_vs_oracle_full_size(B, H, W):
    import cartseg as cs
    from oracle import postproc_oracle as P
    g = torch.Generator().manual_seed(B * 7 + H)
    probs = torch.sigmoid(3.0 * torch.randn(B, H, W, generator=g))
    probs[0, 0, :5] = torch.tensor([0.5, 0.0, 1.0, 0.49999997, 0.50000006])
    if B > 1:
        probs[1] = 0.5                                                   # constant image: all confidences equal
    mask, fg_area, fg_conf, mean_ent = cs.pseudo_label_qc(probs.cuda(), 0.5, mask_value=255)
    mask, fg_area, fg_conf, mean_ent = mask.cpu().numpy(), fg_area.cpu(), fg_conf.cpu(), mean_ent.cpu()
    for b in range(B):
        pred01, a, c, e = P.qc_scores(probs[b].numpy(), 0.5)
        assert np.array_equal(mask[b], pred01 * 255), b
        assert float(fg_area[b]) == a, b
        assert float(fg_conf[b]) == c, b
        assert abs(float(mean_ent[b]) - e) <= 1e-6, b
    assert cs.should_accept(0.2, 0.9, 0.1) and not cs.should_accept(0.7, 0.9, 0.1)
