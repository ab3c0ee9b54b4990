"""Pins oracle/postproc_oracle.py against the reference's own functions (clean_masks.py / remove_blops.py run with
OpenCV, the QC expressions of create_pseudo_labels_gpu.py) — vectors in tests/golden/postproc.npz."""
import os

import numpy as np

from oracle import postproc_oracle as P


def postproc_golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "postproc.npz")))


def test_mask_cleanup_bit_exact_vs_reference(golden_dir):
    g = postproc_golden(golden_dir)
    names = sorted(k[:-5] for k in g if k.endswith("_mask"))
    assert len(names) >= 14
    for n in names:
        m = g[n + "_mask"]
        assert np.array_equal(P.clean_mask(m), g[n + "_clean"]), n
        want = g[n + "_largest"]
        got = P.clean_mask_largest_component(m)
        if want.max() <= 1:                 # remove_blops.py:26-27 returns the {0,1} input when nothing is found
            assert not got.any() and not want.any(), n
        else:
            assert np.array_equal(got, want), n


def test_qc_scores_vs_reference(golden_dir):
    g = postproc_golden(golden_dir)
    for n in ("qc_a", "qc_b", "qc_c"):
        z = g[n + "_logits"]
        probs = P.ensemble_probs([z[0], z[1]], [0.7, 0.3])
        np.testing.assert_allclose(probs, g[n + "_probs"], rtol=0, atol=1.2e-7)
        H, W = probs.shape
        pred01, fg_area, fg_conf, mean_ent = P.qc_scores(g[n + "_probs"])
        want = np.unpackbits(g[n + "_pred01"])[:H * W].reshape(H, W)
        assert np.array_equal(pred01, want), n
        assert fg_area == float(g[n + "_fg_area"]), n
        assert fg_conf == float(g[n + "_fg_conf"]), n                     # median: exact
        assert abs(mean_ent - float(g[n + "_mean_ent"])) <= 1e-6, n


def test_should_accept_thresholds():
    assert P.should_accept(0.2, 0.9, 0.1)
    assert not P.should_accept(0.004, 0.9, 0.1)       # create_pseudo_labels_gpu.py:58-59,142
    assert not P.should_accept(0.61, 0.9, 0.1)
    assert not P.should_accept(0.2, 0.64, 0.1)        # :60,143
    assert not P.should_accept(0.2, 0.9, 0.36)        # :61,144
    assert not P.should_accept(0.2, 0.9, 0.1, edge_hit=0.05)
