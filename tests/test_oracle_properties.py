"""Property tests of the CPU oracles against independent implementations (brute force, scipy, OpenCV when present):
the oracle is what the GPU kernels are compared with at sizes the golden vectors do not cover, so it is itself
checked on randomised small cases here (hypothesis, derandomised so that the suite is reproducible)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import abl_oracle as A
from oracle import edt_oracle as E
from oracle import postproc_oracle as P
from oracle import preproc_oracle as R

SETTINGS = dict(max_examples=40, deadline=None, derandomize=True)


def _mask(draw_seed, h, w, p):
    rng = np.random.Generator(np.random.PCG64(draw_seed))
    return rng.random((h, w)) < p


@settings(**SETTINGS)
@given(st.integers(0, 10 ** 6), st.integers(1, 14), st.integers(1, 14), st.sampled_from([0.05, 0.3, 0.5, 0.8, 0.97]))
def test_edt_squared_equals_brute_force(seed, h, w, p):
    m = _mask(seed, h, w, p)
    if m.all():
        m[0, 0] = False                                   # edt_squared requires at least one zero pixel
    got = E.edt_squared(m)
    zy, zx = np.nonzero(~m)
    yy, xx = np.mgrid[0:h, 0:w]
    d2 = ((yy[..., None] - zy) ** 2 + (xx[..., None] - zx) ** 2).min(-1)
    assert np.array_equal(got, np.where(m, d2, 0))


@settings(**SETTINGS)
@given(st.integers(0, 10 ** 6), st.integers(2, 20), st.integers(2, 20), st.sampled_from([0.0, 0.1, 0.5, 0.9, 1.0]))
def test_sdf_and_abl_distance_channels_equal_scipy(seed, h, w, p):
    ndi = pytest.importorskip("scipy.ndimage")
    m = _mask(seed, h, w, p)
    want = np.zeros((h, w), np.float32)
    if m.any() and (~m).any():
        want = (ndi.distance_transform_edt(~m) - ndi.distance_transform_edt(m)).astype(np.float32)
    assert np.array_equal(E.sdf_of_mask(m).view(np.uint32), want.view(np.uint32))
    for ch in (0, 1):                                    # abl.py:16-24,168-169 through scipy, incl. the no-zero case
        pos = ~m if ch == 0 else m
        res = np.zeros((h, w), np.int32)
        if pos.any():
            neg = ~pos
            res[...] = ndi.distance_transform_edt(neg) * neg - (ndi.distance_transform_edt(pos) - 1) * pos
        assert np.array_equal(A.one_hot_dist_channel(m, ch), np.maximum(-res, 0).astype(np.float32))


@settings(**SETTINGS)
@given(st.integers(0, 10 ** 6), st.integers(1, 24), st.integers(1, 24), st.sampled_from([0.0, 0.2, 0.45, 0.6, 0.9, 1.0]))
def test_mask_cleanup_equals_opencv(seed, h, w, p):
    cv2 = pytest.importorskip("cv2")
    m = (_mask(seed, h, w, p) * 255).astype(np.uint8)
    # clean_masks.py:12-32 with OpenCV
    _, binary = cv2.threshold(m, 127, 255, cv2.THRESH_BINARY)
    filled = binary.copy()
    cv2.floodFill(filled, np.zeros((h + 2, w + 2), np.uint8), (0, 0), 255)
    clean = cv2.bitwise_or(binary, cv2.bitwise_not(filled))
    n, labels, stats, _ = cv2.connectedComponentsWithStats(clean, connectivity=8)
    want = clean if n <= 1 else np.where(labels == 1 + np.argmax(stats[1:, cv2.CC_STAT_AREA]), 255, 0).astype(np.uint8)
    assert np.array_equal(P.clean_mask(m), want)
    # remove_blops.py:14-33
    n, labels, stats, _ = cv2.connectedComponentsWithStats((m > 0).astype(np.uint8), connectivity=8)
    want = np.zeros_like(m) if n <= 1 else (labels == 1 + np.argmax(stats[1:, cv2.CC_STAT_AREA])).astype(np.uint8) * 255
    assert np.array_equal(P.clean_mask_largest_component(m), want)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(3, 90), st.integers(3, 90), st.integers(2, 64), st.integers(2, 64))
def test_resize_equals_opencv(seed, h, w, dh, dw):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(R.resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    m = rng.integers(0, 2, (h, w), dtype=np.uint8) * 255
    assert np.array_equal(R.resize_nearest_u8(m, dh, dw), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST))


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.integers(0, 10 ** 6), st.integers(2, 600))
def test_qc_median_rule(seed, n):
    """np.median on float32: middle element, or the float32 mean of the two middle ones — what cs_pseudo_qc selects."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.random(n).astype(np.float32)
    conf = np.abs(p - np.float32(0.5)) * np.float32(2.0)
    srt = np.sort(conf)
    want = srt[n // 2] if n % 2 else np.float32((srt[n // 2 - 1] + srt[n // 2]) * np.float32(0.5))
    _, _, fg_conf, _ = P.qc_scores(p.reshape(1, n))
    assert np.float32(fg_conf) == want
