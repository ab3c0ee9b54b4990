"""GPU parity of the input side (cs_preproc_images / cs_preproc_masks): letterbox + bilinear resize bit-exact against
OpenCV-produced golden vectors (tests/golden/preproc.npz) and the CPU oracle; normalisation to 1 float32 ulp."""
import numpy as np
import pytest
import torch

from gpu_util import GOLDEN
from test_oracle_preproc import preproc_golden

pytestmark = pytest.mark.gpu


def _as_u8(x_norm: torch.Tensor) -> np.ndarray:
    """mean 0 / std 1 output * 255 -> the resized uint8 image, HWC"""
    return torch.round(x_norm * 255.0).to(torch.uint8).permute(1, 2, 0).cpu().numpy()


def test_letterbox_resize_bit_exact_vs_opencv_golden():
    import cartseg as cs
    cases = preproc_golden(GOLDEN)
    for S in (56, 224):
        sel = [c for c in cases if S in c[4]]
        imgs = [torch.from_numpy(c[1]).cuda() for c in sel]
        out = cs.letterbox_resize_normalize(imgs, S, mean=(0, 0, 0), std=(1, 1, 1))       # one launch, mixed sizes
        assert out.shape == (len(sel), 3, S, S) and out.dtype == torch.float32
        for b, c in enumerate(sel):
            assert np.array_equal(_as_u8(out[b]), c[4][S]), (c[0], S)
        masks = [torch.from_numpy(c[2]).cuda() for c in sel]
        mo = cs.resize_masks(masks, S)
        for b, c in enumerate(sel):
            assert np.array_equal(mo[b, 0].cpu().numpy() == 1.0, c[5][S]), (c[0], S)
            assert set(np.unique(mo[b].cpu().numpy())) <= {0.0, 1.0}


@pytest.mark.parametrize("S", [224, 512])
def test_full_pipeline_vs_oracle(S):
    import cartseg as cs
    from oracle import preproc_oracle as R
    rng = np.random.Generator(np.random.PCG64(S))
    shapes = [(480, 640), (720, 1280), (2 * S, int(2 * S / 1.2)), (333, 217), (S, S), (64, 48)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    out = cs.letterbox_resize_normalize(dev, S, bgr=True).cpu().numpy()                  # ImageNet statistics
    for b, im in enumerate(imgs):
        want = R.preprocess_image(im, S, bgr=True)
        np.testing.assert_allclose(out[b], want, rtol=0, atol=2.4e-7 * 3, err_msg=str(shapes[b]))
    out0 = cs.letterbox_resize_normalize(dev, S, mean=(0, 0, 0), std=(1, 1, 1), letterbox=False)
    for b, im in enumerate(imgs):
        assert np.array_equal(_as_u8(out0[b]), R.resize_linear_u8(im, S, S)), shapes[b]
    masks = [(rng.random((h, w)) < 0.3).astype(np.uint8) * 255 for h, w in shapes]
    mo = cs.resize_masks([torch.from_numpy(m).cuda() for m in masks], S).cpu().numpy()
    for b, m in enumerate(masks):
        assert np.array_equal(mo[b], R.preprocess_mask(m, S)), shapes[b]


def test_preprocessed_batch_feeds_the_model():
    import cartseg as cs
    rng = np.random.Generator(np.random.PCG64(1))
    imgs = [torch.from_numpy(rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)).cuda() for _ in range(2)]
    x = cs.letterbox_resize_normalize(imgs, 64)
    torch.manual_seed(0)
    net = cs.UNet().cuda().eval()
    with torch.no_grad():
        z = net(x)
    assert z.shape == (2, 1, 64, 64) and torch.isfinite(z).all()
