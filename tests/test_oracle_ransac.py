"""Pins the geometric-verification restatement (oracle/ransac_oracle.c: cv::estimateAffinePartial2D's RANSAC loop as the
reference calls it, image_utils.rs:45-60) against cv2: committed golden masks + live randomised trials."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_ransac_masks_equal_golden():
    g = np.load(os.path.join(GOLD, "ransac.npz"))
    for i in range(int(g["n_sets"][0])):
        rating, mask, _, _ = oracle.ransac_affine_partial(g[f"from{i}"], g[f"to{i}"])
        assert np.array_equal(mask, g[f"mask{i}"]), f"set {i}"
        assert rating == int(g[f"mask{i}"].sum())


def test_degenerate_sizes():
    z = np.zeros((0, 2), np.float32)
    assert oracle.ransac_affine_partial(z, z)[0] == 0
    one = np.ones((1, 2), np.float32)
    assert oracle.ransac_affine_partial(one, one)[0] == 0          # fewer than modelPoints -> no model, no inliers
    two = np.array([[0, 0], [10, 0]], np.float32)
    assert oracle.ransac_affine_partial(two, two + 5)[0] == 2      # count == modelPoints -> every point is an inlier


def test_update_num_iters_table():
    L = oracle.lib()
    assert L.ransac_update_num_iters(0.99, 0.0, 2, 2000) == 0      # all inliers: denom = log(0) -> -inf, num/denom = 0
    assert L.ransac_update_num_iters(0.99, 1.0, 2, 2000) == 2000   # no inliers: unchanged
    assert L.ransac_update_num_iters(0.99, 0.5, 2, 2000) == 16
    assert L.ransac_update_num_iters(0.99, 0.9, 2, 2000) == 458
    assert L.ransac_update_num_iters(0.99, 0.9, 2, 100) == 100


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_ransac_equals_cv2_live():
    rng = np.random.default_rng(5)
    for trial in range(120):
        n = int(rng.integers(2, 600))
        fr = rng.uniform(0, 2000, (n, 2)).astype(np.float32)
        s, a = rng.uniform(0.9, 1.1), rng.uniform(-0.05, 0.05)
        R = np.array([[s * np.cos(a), -s * np.sin(a)], [s * np.sin(a), s * np.cos(a)]])
        to = (fr @ R.T + rng.uniform(-20, 20, 2)).astype(np.float32)
        out = rng.random(n) > rng.choice([0.0, 0.05, 0.3, 0.9])
        to[out] = rng.uniform(0, 2000, (int(out.sum()), 2)).astype(np.float32)
        to += rng.normal(0, 2.1, to.shape).astype(np.float32)          # many residuals near the 3 px threshold
        if trial % 10 == 0 and n > 4:
            fr[1], to[1] = fr[0], to[0]                                  # duplicate correspondence (degenerate sample)
        _, inl = cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ransacReprojThreshold=3.0, maxIters=2000, confidence=0.99,
                                             refineIters=10)
        want = inl.ravel() if inl is not None else np.zeros(n, np.uint8)
        rating, mask, _, _ = oracle.ransac_affine_partial(fr, to)
        assert np.array_equal(mask, want), f"trial {trial} n={n}"
