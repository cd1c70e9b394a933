"""K8/K9 parity: GPU brute-force Hamming k-NN (through the C ABI) == the oracle, bit-exact indices and distances.

Oracle = oracle/bf_oracle.c (pinned against cv2.BFMatcher in tests/test_oracle_bf.py).  Covers the edge cases the
domain has: ties across pages (planted duplicates), exact hits (distance 0 -> no vote), nt < k (padded rows),
ragged sizes, pool splits (small nq x large nt), k < 30, empty inputs.
"""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


def _check(ctx, nq, nt, k, seed=0):
    pool = synth.hamming_pool(nt, seed=7 + seed, dup_frac=0.02) if nt else np.zeros((0, 32), np.uint8)
    q = synth.hamming_queries(pool, nq, seed=9 + seed)
    gi, gd = ctx.bf_knn_hamming(q, pool, k)
    oi, od = oracle.bf_knn_hamming(q, pool, k)
    assert np.array_equal(gd, od), f"distances differ nq={nq} nt={nt} k={k}"
    assert np.array_equal(gi, oi), f"indices differ nq={nq} nt={nt} k={k}"


@pytest.mark.parametrize("nq,nt,k", [
    (1, 1, 30), (1, 29, 30), (3, 30, 30), (7, 31, 30), (33, 257, 30), (512, 1000, 30), (513, 2049, 30),
    (2092, 4111, 30), (100, 20000, 30), (17, 100000, 30), (1500, 33333, 30), (300, 5000, 1), (300, 5000, 7),
    (300, 5000, 32), (64, 255, 16), (5, 0, 30),
])
def test_knn_matches_oracle(ctx, nq, nt, k):
    _check(ctx, nq, nt, k)


def test_knn_all_identical_pool(ctx):
    # every pooled descriptor identical: the k-boundary is decided by index order alone
    pool = np.tile(np.arange(32, dtype=np.uint8), (500, 1))
    q = np.tile(np.arange(32, dtype=np.uint8), (40, 1))
    q[1::2, 0] ^= 0xFF
    gi, gd = ctx.bf_knn_hamming(q, pool, 30)
    assert np.array_equal(gi, np.tile(np.arange(30, dtype=np.int32), (40, 1)))
    assert np.array_equal(gd[0::2], np.zeros((20, 30), np.int32))
    assert np.array_equal(gd[1::2], np.full((20, 30), 8, np.int32))


def test_knn_extreme_distances(ctx):
    pool = np.zeros((100, 32), np.uint8)
    pool[50:] = 0xFF
    q = np.zeros((2, 32), np.uint8)
    q[1] = 0xFF
    gi, gd = ctx.bf_knn_hamming(q, pool, 30)
    oi, od = oracle.bf_knn_hamming(q, pool, 30)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    gi, gd = ctx.bf_knn_hamming(q, pool, 32)       # k spans both halves? no: 32 < 50, all distance 0
    assert gd.max() == 0
    pool2 = pool[40:60]                             # 10 zeros + 10 ones: rows contain distance 0 and 256
    gi, gd = ctx.bf_knn_hamming(q, pool2, 20)
    oi, od = oracle.bf_knn_hamming(q, pool2, 20)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od) and gd.max() == 256


def test_match_descriptors_vote_matches_oracle(ctx):
    import slideo_b200
    rng = np.random.default_rng(5)
    pages = [synth.hamming_pool(int(n), seed=100 + i, dup_frac=0.0) for i, n in enumerate(rng.integers(0, 400, 12))]
    pages[3] = pages[2].copy()                      # identical pages: ties across pages
    pages[5] = np.zeros((0, 32), np.uint8)          # a page without keypoints
    pool = np.concatenate(pages)
    frames = [synth.hamming_queries(pool, int(n), seed=200 + i, near_frac=0.7) for i, n in enumerate([300, 0, 1, 257, 999])]
    fo = np.zeros(len(frames) + 1, np.int32)
    fo[1:] = np.cumsum([len(f) for f in frames])
    with slideo_b200.Context(slideo_b200.default_config(keep_matches=1, max_batch=2)) as c:
        for p in pages:
            c.add_page_descriptors(p)
        c.finalize_pool()
        assert c.pool_info() == (len(pool), len(pages))
        res = c.match_descriptors(np.concatenate(frames), fo)
        for i, f in enumerate(frames):
            best, votes, _ = oracle.match_frame(f, pages)
            assert (res[i, 0], res[i, 1], res[i, 2]) == (best, votes, len(f)), f"frame {i}"
            m = c.get_matches(i)
            assert m.shape == (len(f), 30)
            if len(f):
                oi, od = oracle.bf_knn_hamming(f, pool, 30)
                offs = np.zeros(len(pages) + 1, np.int64)
                offs[1:] = np.cumsum([len(p) for p in pages])
                assert np.array_equal(offs[m["source"]] + m["train_idx"], oi)
                assert np.array_equal(m["distance"].astype(np.int32), od)
                assert np.array_equal(m["query_idx"], np.tile(np.arange(len(f), dtype=np.int32)[:, None], (1, 30)))


def test_knn_full_size_properties(ctx):
    """BASELINE config-2 geometry (2k queries x 100k pool): size-independent properties instead of the slow oracle."""
    pool = synth.hamming_pool(100_000, seed=11)
    q = synth.hamming_queries(pool, 2048, seed=12)
    gi, gd = ctx.bf_knn_hamming(q, pool, 30)
    # rows ascending by (distance, index), indices unique and in range
    key = gd.astype(np.int64) * (1 << 32) + gi
    assert (np.diff(key, axis=1) > 0).all()
    assert gi.min() >= 0 and gi.max() < len(pool)
    # reported distances are the true Hamming distances
    x = np.bitwise_xor(q[:, None, :], pool[gi])
    true = np.unpackbits(x, axis=2).sum(axis=2)
    assert np.array_equal(true, gd)
    # nothing outside the row beats the row's last entry: check a random sample of queries exhaustively
    rng = np.random.default_rng(3)
    for i in rng.integers(0, len(q), 24):
        d = np.unpackbits(np.bitwise_xor(q[i][None, :], pool), axis=1).sum(axis=1)
        order = np.argsort(d, kind="stable")[:30]
        assert np.array_equal(order.astype(np.int32), gi[i])
    # permutation property: reversing the pool order maps indices (ties aside) to the same distances
    gi2, gd2 = ctx.bf_knn_hamming(q, pool[::-1].copy(), 30)
    assert np.array_equal(gd2, gd)
