"""Pins the matcher + vote restatement (oracle/bf_oracle.c) against cv2.BFMatcher golden vectors and cv2 live."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "bf_knn.npz"))


def test_hamming_knn_equals_golden(g):
    idx, dist = oracle.bf_knn_hamming(g["ham_q"], g["ham_pool"], 30)
    assert np.array_equal(idx, g["ham_idx"]) and np.array_equal(dist, g["ham_dist"])
    assert (g["ham_dist"][0, 0] == 0)                      # the planted exact hit


def test_vote_equals_golden(g):
    lens = g["ham_pages_len"]
    offs = np.zeros(len(lens) + 1, np.int32)
    offs[1:] = np.cumsum(lens)
    idx, dist = oracle.bf_knn_hamming(g["ham_q"], g["ham_pool"], 30)
    best, votes, allv = oracle.vote(idx, dist.astype(np.float32), offs)
    assert np.array_equal(allv, g["ham_votes"])
    assert best == int(np.argmax(g["ham_votes"])) and votes == int(g["ham_votes"].max())
    # the exact-hit query (best distance 0) casts no vote: lib.rs:275 `d < 0 * 1.05` is never true
    b0, v0, a0 = oracle.vote(idx[:1], dist[:1].astype(np.float32), offs)
    assert (b0, v0, int(a0.sum())) == (-1, 0, 0)


def test_l2_knn_equals_golden(g):
    idx, dist = oracle.bf_knn_l2(g["l2_q"], g["l2_t"], 30)
    assert np.array_equal(idx, g["l2_idx"])
    assert np.array_equal(dist.view(np.uint32), g["l2_dist"].view(np.uint32))


def test_ratio_rule_integer_equivalence():
    """(float)d < (float)best * 1.05f  <=>  20 d < 21 best  for all Hamming distances."""
    d = np.arange(257, dtype=np.float32)[:, None]
    b = np.arange(257, dtype=np.float32)[None, :]
    assert np.array_equal(d < b * np.float32(1.05), 20 * d.astype(np.int64) < 21 * b.astype(np.int64))


def test_padding_when_pool_smaller_than_k():
    rng = np.random.default_rng(1)
    q = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    idx, dist = oracle.bf_knn_hamming(q, t, 30)
    assert (idx[:, 5:] == -1).all() and (dist[:, 5:] == -1).all() and (idx[:, :5] >= 0).all()
    idx, dist = oracle.bf_knn_hamming(q, np.zeros((0, 32), np.uint8), 30)
    assert (idx == -1).all()


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_hamming_knn_equals_cv2_live():
    from oracle import cv2_oracle as co
    import synth
    pages = [synth.hamming_pool(n, seed=50 + i, dup_frac=0.1) for i, n in enumerate((300, 41, 500))]
    pool = np.concatenate(pages)
    q = synth.hamming_queries(pool, 200, seed=51)
    ci, cd = co.bf_knn_hamming(q, pages, 30)
    oi, od = oracle.bf_knn_hamming(q, pool, 30)
    assert np.array_equal(ci, oi) and np.array_equal(cd, od)
    b, v, allv = co.match_frame_bf(q, pages)
    ob, ov, oall = oracle.match_frame(q, pages)
    assert (b, v) == (ob, ov) and np.array_equal(allv, oall)
