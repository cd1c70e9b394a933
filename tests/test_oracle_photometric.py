"""Pins oracle/photometric.py (warpAffine nearest inverse map + LM refinement + the similarity gate, lib.rs:335-389) against cv2."""
import numpy as np
import pytest

import oracle
from oracle import photometric as ph
import synth

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

pytestmark = pytest.mark.skipif(cv2 is None, reason="cv2 not importable")


def test_warp_equals_cv2():
    rng = np.random.default_rng(0)
    fr = synth.make_frame(3, 50)
    for _ in range(3):
        s, a = rng.uniform(0.9, 1.0), rng.uniform(-0.03, 0.03)
        m = np.array([[s * np.cos(a), -s * np.sin(a), rng.uniform(-30, 30)], [s * np.sin(a), s * np.cos(a), rng.uniform(-30, 30)]])
        ref = cv2.warpAffine(fr, m, (2001, 1125), flags=cv2.WARP_INVERSE_MAP, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(ph.warp_affine_inverse_nearest(fr, m, 2001, 1125), ref)


def test_refined_matrix_close_to_cv2():
    rng = np.random.default_rng(3)
    for _ in range(12):
        n = int(rng.integers(50, 1500))
        fr = rng.uniform(0, 2000, (n, 2)).astype(np.float32)
        s, a = rng.uniform(0.9, 1.1), rng.uniform(-0.05, 0.05)
        r = np.array([[s * np.cos(a), -s * np.sin(a)], [s * np.sin(a), s * np.cos(a)]])
        to = (fr @ r.T + rng.uniform(-20, 20, 2)).astype(np.float32)
        out = rng.random(n) > rng.choice([0.3, 0.6, 0.9])
        to[out] = rng.uniform(0, 2000, (int(out.sum()), 2)).astype(np.float32)
        to += rng.normal(0, 1.0, to.shape).astype(np.float32)
        m_ref, inl = cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ransacReprojThreshold=3.0, maxIters=2000, confidence=0.99,
                                                  refineIters=10)
        m, mask = ph.refined_matrix(fr, to)
        assert np.array_equal(mask, inl.ravel())
        assert np.abs(m - m_ref).max() < 1e-9          # tolerance: the 4x4 solver is not restated bit for bit


def test_full_tail_on_a_synthetic_frame():
    """frame -> votes -> RANSAC -> warp -> similarity with cv2 doing warp/resize/norm, against the restatement."""
    pages = [synth.make_page(p) for p in range(3)]
    feats = [oracle.orb_detect_and_compute(p) for p in pages]
    pool = np.concatenate([f[2] for f in feats])
    pts = np.concatenate([f[1][:, :2] for f in feats])
    offs = np.zeros(4, np.int32)
    offs[1:] = np.cumsum([len(f[2]) for f in feats])
    frame = synth.make_frame(1, 3, pages)
    ki, kf, d = oracle.orb_detect_and_compute(oracle.gray_from_bgr(frame))
    idx, dist = oracle.bf_knn_hamming(d, pool, 30)
    res = ph.decide_frame(idx, dist.astype(np.float32), offs, kf[:, :2], pts, frame, pages)
    assert res["image"] == 1 and len(res["rated"]) == 1
    p, sim = res["rated"][0]
    m = res["matrices"][p]
    proj = cv2.warpAffine(frame, m, (2001, 1125), flags=cv2.WARP_INVERSE_MAP, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    small = cv2.resize(proj, (461, 259), interpolation=cv2.INTER_AREA)
    page_small = cv2.resize(cv2.cvtColor(pages[p], cv2.COLOR_GRAY2BGR), (461, 259), interpolation=cv2.INTER_AREA)
    err = cv2.norm(small, page_small, cv2.NORM_L2)
    want = np.float32(1.0) - np.float32(err) / np.sqrt(np.float32(255.0 * 255.0 * 3.0) * np.float32(259 * 461), dtype=np.float32)
    assert sim == want and sim > 0.5
