"""K10 parity: GPU L2 k-NN (bf16 tcgen05 GEMM, through the C ABI) vs the oracle (oracle/bf_oracle.c, pinned against
cv2.BFMatcher(NORM_L2) in tests/test_oracle_bf.py).

Integer-valued descriptors (what cv2.SIFT emits): indices AND distances bit-exact.
General float descriptors: distances within 1e-4 relative of the bf16-rounded inputs' true distances is NOT claimed --
bf16 rounding of the inputs dominates; the test bounds the error against the oracle run on the rounded inputs
(tolerance 1e-4 relative, the figure BASELINE.json's north_star states for SIFT L2)."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sift_like(n, seed, dup_from=None):
    """Integer-valued rows in 0..255 with row norm ~512 like cv2 SIFT descriptors."""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-6) * 512.0
    x = np.minimum(np.rint(x), 255).astype(np.float32)
    if dup_from is not None and len(dup_from) and n:
        src = rng.integers(0, len(dup_from), max(1, n // 20))
        dst = rng.integers(0, n, len(src))
        x[dst] = dup_from[src]
    return x


def _check_exact(ctx, nq, nt, k, seed=0):
    t = sift_like(nt, 10 + seed)
    if nt > 40:
        t[nt // 2: nt // 2 + 10] = t[:10]            # exact duplicates inside the pool -> ties decided by index
    q = sift_like(nq, 20 + seed, dup_from=t)          # some queries are exact pool rows -> distance 0
    gi, gd = ctx.bf_knn_l2(q, t, k)
    oi, od = oracle.bf_knn_l2(q, t, k)
    assert np.array_equal(gi, oi), f"indices differ nq={nq} nt={nt} k={k}"
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32)), f"distances differ nq={nq} nt={nt} k={k}"


@pytest.mark.parametrize("nq,nt,k", [
    (1, 1, 30), (5, 29, 30), (7, 255, 30), (128, 256, 30), (129, 257, 30), (300, 1000, 30), (1000, 5000, 30),
    (64, 20000, 30), (2500, 3000, 1), (500, 3000, 7), (500, 3000, 32), (3, 0, 30), (20000, 777, 30),
])
def test_l2_knn_matches_oracle_exactly_on_integer_descriptors(ctx, nq, nt, k):
    _check_exact(ctx, nq, nt, k)


def test_l2_golden_from_cv2(ctx):
    g = np.load(os.path.join(GOLD, "bf_knn.npz"))
    gi, gd = ctx.bf_knn_l2(g["l2_q"], g["l2_t"], 30)
    assert np.array_equal(gi, g["l2_idx"])
    assert np.array_equal(gd.view(np.uint32), g["l2_dist"].view(np.uint32))


def test_l2_float_descriptors_within_tolerance(ctx):
    rng = np.random.default_rng(4)
    t = rng.normal(0, 1, (4000, 128)).astype(np.float32)
    q = (t[rng.integers(0, 4000, 200)] + rng.normal(0, 0.3, (200, 128))).astype(np.float32)
    import torch
    tb = torch.from_numpy(t).to(torch.bfloat16).to(torch.float32).numpy()     # the rounding the kernel applies
    qb = torch.from_numpy(q).to(torch.bfloat16).to(torch.float32).numpy()
    gi, gd = ctx.bf_knn_l2(q, t, 30)
    d_true = np.sqrt(((qb[:, None, :].astype(np.float64) - tb[gi].astype(np.float64)) ** 2).sum(-1))
    assert np.allclose(gd, d_true, rtol=1e-4, atol=1e-5)                      # tolerance: 1e-4 relative (north_star)
    oi, od = oracle.bf_knn_l2(qb, tb, 30)
    assert np.allclose(gd, od, rtol=1e-4, atol=1e-5)
    assert (np.diff(gd, axis=1) >= 0).all()
    agree = (gi == oi).mean()
    assert agree > 0.99                                                        # index flips only between near-equal distances


def test_sift128_pool_vote_matches_oracle():
    import slideo_b200
    pages = [sift_like(n, 50 + i) for i, n in enumerate((300, 0, 450, 280))]
    pages[3][:40] = pages[0][:40]
    pool = np.concatenate(pages)
    frames = [sift_like(n, 70 + i, dup_from=pool) for i, n in enumerate((200, 0, 333))]
    fo = np.zeros(len(frames) + 1, np.int32)
    fo[1:] = np.cumsum([len(f) for f in frames])
    offs = np.zeros(len(pages) + 1, np.int32)
    offs[1:] = np.cumsum([len(p) for p in pages])
    cfg = slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, keep_matches=1, max_batch=2)
    with slideo_b200.Context(cfg) as c:
        for p in pages:
            c.add_page_descriptors(p)
        c.finalize_pool()
        res = c.match_descriptors(np.concatenate(frames), fo)
        for i, f in enumerate(frames):
            if len(f) == 0:
                assert tuple(res[i]) == (-1, 0, 0)
                continue
            oi, od = oracle.bf_knn_l2(f, pool, 30)
            best, votes, _ = oracle.vote(oi, od, offs)
            assert tuple(res[i]) == (best, votes, len(f)), f"frame {i}"
            m = c.get_matches(i)
            assert np.array_equal(offs[m["source"]] + m["train_idx"], oi)
            assert np.array_equal(m["distance"].view(np.uint32), od.view(np.uint32))


def test_l2_full_size_properties(ctx):
    """BASELINE config-4-like geometry (2k queries x 200k pool): order, range and exactness of reported distances."""
    t = sift_like(200_000, 90)
    q = sift_like(2048, 91, dup_from=t)
    gi, gd = ctx.bf_knn_l2(q, t, 30)
    assert gi.min() >= 0 and gi.max() < len(t)
    assert (np.diff(gd, axis=1) >= 0).all()
    d2 = ((q[:, None, :] - t[gi]) ** 2).sum(-1)
    assert np.array_equal(gd, np.sqrt(d2).astype(np.float32))
    rng = np.random.default_rng(1)
    for i in rng.integers(0, len(q), 16):
        d = np.sqrt(((q[i][None, :] - t) ** 2).sum(-1)).astype(np.float32)
        order = np.lexsort((np.arange(len(t)), d))[:30]
        assert np.array_equal(order.astype(np.int32), gi[i])


def _exact_knn_rows(q, t, k, rows):
    """(distance, index)-ordered k nearest pooled rows of the given query rows: exact integer arithmetic in numpy (the oracle's
    definition, oracle/bf_oracle.c, restated for a handful of rows of a pool too large for the C loop in a test)."""
    out_i, out_d = [], []
    tt = t.astype(np.int64)
    for r in rows:
        d2 = ((tt - q[r].astype(np.int64)[None, :]) ** 2).sum(-1)
        order = np.lexsort((np.arange(len(t)), d2))[:k]
        out_i.append(order.astype(np.int32))
        out_d.append(np.sqrt(d2[order].astype(np.float32)))
    return np.stack(out_i), np.stack(out_d)


def test_l2_long_work_items_take_the_threshold_pre_pass(ctx):
    """Work items of >= 192 pool tiles run the threshold pre-pass (knn_l2.cu, l2_npre) and start selecting against a bound derived
    from group minima; 149 query tiles also put one tile into the split last wave (no pre-pass there).  Exactness is checked row by
    row against integer arithmetic on 64 rows (first, last, split-wave rows, random rows); every row is checked for order, range and
    for reporting the exact distance of the index it names; duplicates inside the pool decide ties by index."""
    import torch
    nq, nt, k = 149 * 128 - 5, 50_000, 30
    t = sift_like(nt, 300)
    t[40_000:40_016] = t[100:116]                 # exact duplicates far apart: ties broken by index
    q = sift_like(nq, 301, dup_from=t)
    gi, gd = ctx.bf_knn_l2(q, t, k)
    assert gi.min() >= 0 and gi.max() < nt
    assert (np.diff(gd, axis=1) >= 0).all()
    # the reported distance is the exact distance of the reported index, for every row (float64 on the GPU box's torch, integers)
    tq, tt = torch.from_numpy(q).cuda().double(), torch.from_numpy(t).cuda().double()
    for a in range(0, nq, 2048):
        sel = torch.from_numpy(gi[a:a + 2048].astype(np.int64)).cuda()
        d2 = ((tq[a:a + 2048, None, :] - tt[sel]) ** 2).sum(-1)
        assert np.array_equal(gd[a:a + 2048], np.sqrt(d2.cpu().numpy().astype(np.float32)))
    # no row misses a neighbour: the k-th reported distance equals the true k-th smallest distance (all rows, chunked on the GPU)
    tn = (tt * tt).sum(1)
    for a in range(0, nq, 1024):
        d2 = (tq[a:a + 1024] ** 2).sum(1, keepdim=True) + tn[None, :] - 2.0 * tq[a:a + 1024] @ tt.T
        kth = torch.topk(d2, k, dim=1, largest=False).values[:, -1].cpu().numpy()
        assert np.array_equal(gd[a:a + 1024, -1], np.sqrt(kth.astype(np.float32)))
    rng = np.random.default_rng(5)
    rows = np.unique(np.concatenate([[0, 1, 127, 128, nq - 1, nq - 2, 148 * 128, 148 * 128 + 60], rng.integers(0, nq, 56)]))
    oi, od = _exact_knn_rows(q, t, k, rows)
    assert np.array_equal(gi[rows], oi)
    assert np.array_equal(gd[rows].view(np.uint32), od.view(np.uint32))


def test_l2_pre_pass_with_float_descriptors(ctx):
    """The same long work items with general float rows: the pre-pass bound must hold for inexact accumulations too (tolerance as in
    test_l2_float_descriptors_within_tolerance: 1e-4 relative against the bf16-rounded inputs, the figure north_star states)."""
    import torch
    rng = np.random.default_rng(9)
    nq, nt, k = 148 * 128, 50_000, 30
    t = rng.normal(0, 1, (nt, 128)).astype(np.float32)
    q = (t[rng.integers(0, nt, nq)] + rng.normal(0, 0.3, (nq, 128))).astype(np.float32)
    gi, gd = ctx.bf_knn_l2(q, t, k)
    assert (np.diff(gd, axis=1) >= 0).all()
    tq = torch.from_numpy(q).cuda().to(torch.bfloat16).double()      # the rounding the kernel applies
    tt = torch.from_numpy(t).cuda().to(torch.bfloat16).double()
    tn = (tt * tt).sum(1)
    for a in range(0, nq, 1024):
        d2 = ((tq[a:a + 1024] ** 2).sum(1, keepdim=True) + tn[None, :] - 2.0 * tq[a:a + 1024] @ tt.T).clamp_(min=0)
        true = torch.topk(d2, k, dim=1, largest=False).values.sqrt().cpu().numpy()
        assert np.allclose(gd[a:a + 1024], true, rtol=1e-4, atol=1e-5)
