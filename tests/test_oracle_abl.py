"""Pins oracle/abl_oracle.py against the reference's own ABL class (tests/golden/abl.npz was produced by
oracle/make_golden.py running src/training/losses/abl.py unchanged on the CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import abl_oracle as A


def abl_golden_cases(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "abl.npz")))
    for n in sorted(k[:-6] for k in g if k.endswith("_shape")):
        shp = tuple(int(v) for v in g[n + "_shape"])
        B, _, H, W = shp
        unpack = lambda bits, s: np.unpackbits(bits)[:int(np.prod(s))].reshape(s).astype(bool)  # noqa: E731
        yield n, dict(
            logits=torch.from_numpy(g[n + "_logits"]),
            targets=torch.from_numpy(unpack(g[n + "_targets"], shp).astype(np.float32)),
            pred_boundary=unpack(g[n + "_pred_boundary"], (B, H, W)),
            gt_boundary=unpack(g[n + "_gt_boundary"], (B, H, W)),
            dist_maps=g[n + "_dist_maps"].astype(np.float32),
            none=bool(g[n + "_none"]), value=g.get(n + "_value"), grad=g.get(n + "_grad"))


def test_abl_oracle_matches_reference(golden_dir):
    seen_none = seen_loss = 0
    for n, c in abl_golden_cases(golden_dir):
        x = c["logits"].clone().requires_grad_(True)
        loss, parts = A.abl_loss(x, c["targets"], return_parts=True)
        assert np.array_equal(parts["pred_boundary"].numpy(), c["pred_boundary"]), n      # bit-exact mask
        if c["none"]:
            assert loss is None, n
            seen_none += 1
            continue
        seen_loss += 1
        B = x.shape[0]
        assert np.array_equal(parts["gt_boundary"], c["gt_boundary"]), n
        # the reference indexes its [2B,H,W] concatenation with the batch index: entries 0..B-1 are what it uses
        assert np.array_equal(parts["dmap"].numpy(), c["dist_maps"][:B]), n
        assert float(loss.detach()) == pytest.approx(float(c["value"]), rel=2e-6, abs=1e-7), n
        loss.backward()
        scale = max(float(np.abs(c["grad"]).max()), 1e-12)
        assert float((x.grad - torch.from_numpy(c["grad"])).abs().max()) <= 1e-5 * scale + 1e-10, n
    assert seen_none >= 1 and seen_loss >= 6


def test_abl_dist_channels_match_scipy_formula():
    """one_hot_dist_channel == max(0, -one_hot2dist(...)) computed with scipy (abl.py:16-24,168-169), including
    scipy's behaviour for inputs without any zero pixel (oracle docstring, behaviour 4)."""
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.Generator(np.random.PCG64(5))
    masks = [rng.random((19, 23)) < p for p in (0.0, 0.02, 0.3, 0.9, 1.0)]
    for gb in masks:
        for ch in (0, 1):
            pos = ~gb if ch == 0 else gb
            res = np.zeros(gb.shape, np.int32)
            if pos.any():
                neg = ~pos
                res[...] = ndi.distance_transform_edt(neg) * neg - (ndi.distance_transform_edt(pos) - 1) * pos
            want = np.maximum(-res, 0).astype(np.float32)
            assert np.array_equal(A.one_hot_dist_channel(gb, ch), want)


def test_eps_ladder_is_the_loop_sequence():
    lad = A.eps_ladder(8)
    e, seq = 1e-5, []
    for _ in range(8):
        seq.append(np.float32(e))
        e *= 1.2
    assert np.array_equal(lad, np.asarray(seq, np.float32))
    big = A.eps_ladder()
    assert big[-1] > 0.93      # KL of two 2-way softmaxes over probabilities is < 0.4622 per direction
