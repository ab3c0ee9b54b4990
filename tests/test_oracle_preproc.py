"""Pins oracle/preproc_oracle.py against the reference's letterbox function and OpenCV's resize
(tests/golden/preproc.npz, produced by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import preproc_oracle as R


def preproc_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "preproc.npz")))
    cases = []
    for name in sorted(k[:-6] for k in g if k.endswith("_image")):
        img = g[name + "_image"]
        H, W = img.shape[:2]
        mask = np.unpackbits(g[name + "_mask"])[:H * W].reshape(H, W).astype(np.uint8) * 255
        sizes = sorted(int(k.rsplit("_", 1)[1]) for k in g if k.startswith(name + "_resized_"))
        cases.append((name, img, mask, int(g[name + "_letterbox_side"]),
                      {S: g[f"{name}_resized_{S}"] for S in sizes},
                      {S: np.unpackbits(g[f"{name}_mask_{S}"])[:S * S].reshape(S, S).astype(bool) for S in sizes}))
    return cases


def test_letterbox_and_bilinear_resize_bit_exact_vs_opencv(golden_dir):
    cases = preproc_golden(golden_dir)
    assert len(cases) >= 6
    for name, img, mask, side, resized, masks in cases:
        lb = R.letterbox(img, 0.1)
        assert lb.shape[0] == lb.shape[1] == side, name
        for S, want in resized.items():
            assert np.array_equal(R.resize_linear_u8(lb, S, S), want), (name, S)
        for S, want in masks.items():
            assert np.array_equal(R.resize_nearest_u8(mask, S, S) > 0, want), (name, S)


def test_resize_matches_opencv_directly_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.Generator(np.random.PCG64(4))
    for (H, W, dh, dw) in [(300, 400, 224, 224), (448, 448, 224, 224), (100, 90, 224, 224), (1080, 1080, 512, 512),
                           (63, 64, 31, 32), (37, 53, 224, 224), (224, 224, 224, 224)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(R.resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        m = rng.integers(0, 2, (H, W), dtype=np.uint8) * 255
        assert np.array_equal(R.resize_nearest_u8(m, dh, dw), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_letterbox_geometry_uses_python_rounding():
    # width 25 -> 2.5 -> round() = 2 (banker's), width 35 -> 3.5 -> 4
    assert R.letterbox_geometry(10, 25)[0] == 29 and R.letterbox_geometry(10, 35)[0] == 43
    L, x0, y0 = R.letterbox_geometry(120, 50)
    assert (L, x0, y0) == (120, 35, 0)


def test_normalize_formula():
    img = np.arange(2 * 2 * 3, dtype=np.uint8).reshape(2, 2, 3) * 20
    out = R.normalize_chw(img, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    ref = (img.astype(np.float64) / 255.0 - np.array([0.485, 0.456, 0.406])) / np.array([0.229, 0.224, 0.225])
    assert out.shape == (3, 2, 2) and out.dtype == np.float32
    np.testing.assert_allclose(out, ref.transpose(2, 0, 1), rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(R.normalize_chw(img, (0, 0, 0), (1, 1, 1)),
                                  (img.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1))
