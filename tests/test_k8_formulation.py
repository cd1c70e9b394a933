"""CPU restatement of the arithmetic K8 v5 relies on (slideo_b200/csrc/knn_hamming5.cu), checked against plain popcount distances.

Test infrastructure only: numpy on 32-bit words, one "lane word" = 32 pooled rows.  What is pinned here without a GPU:
* the carry-save (Harley-Seal) walk over a query's set-bit list gives s = 2 |q & t| + (256 - popc t) as ten bit planes, also when
  the walk stops at the end of the list (whole blocks of 16 entries, or half blocks of 8: the weight-8 carry of a lone half block
  meets no partner), for normal and complemented lists:  d = A + 256 - s  resp.  d = A - 256 + s;
* XORing pool rows and queries with one vector (the pool's majority vector) changes no distance and shortens the lists of biased
  descriptors;
* the dealing of a tile's 128 queries to 16 warps by list length is a bijection and evens out the per-warp walk;
* the query-to-tile map  row = ql * n_tiles + tile  covers a launch's query range exactly once.
The device code itself is compared with the oracle bit for bit in tests/test_gpu_knn.py and tests/test_gpu_stream.py.
"""
import numpy as np
import pytest

ZERO_ROW = 266


def _csa(a, b, c):
    return a ^ b ^ c, (a & b) | (a & c) | (b & c)


def _walk(rows, ptp_planes, lst, nh, gran):
    """k5_scan: rows[b] = lane word of bit row b; ptp_planes[p] = plane p of 256 - popc(t); nh = list length in half blocks."""
    ones, twos, fours, eights, s16, s32, s64, s128 = (np.uint32(ptp_planes[i]) for i in range(1, 9))
    s256 = o_prev = t32a = u64a = np.uint32(0)
    for blk in range(8):
        if gran == 0 or 2 * blk < nh:
            n_in = 16 if (gran != 2 or 2 * blk + 1 < nh) else 8
            x = [rows[lst[blk * 16 + i]] for i in range(n_in)]
            halves = []
            for h in range(n_in // 8):
                y = x[8 * h:8 * h + 8]
                ones, ta = _csa(ones, y[0], y[1]); ones, tb = _csa(ones, y[2], y[3]); twos, fa = _csa(twos, ta, tb)
                ones, ta = _csa(ones, y[4], y[5]); ones, tb = _csa(ones, y[6], y[7]); twos, fb = _csa(twos, ta, tb)
                fours, e = _csa(fours, fa, fb)
                halves.append(e)
            if len(halves) == 2:
                eights, o = _csa(eights, halves[0], halves[1])
            else:
                o, eights = eights & halves[0], eights ^ halves[0]
        else:
            o = np.uint32(0)
        if blk & 1:
            s16, t = _csa(s16, o_prev, o)
            if (blk & 3) == 3:
                s32, u = _csa(s32, t32a, t)
                if blk == 7:
                    s64, v = _csa(s64, u64a, u)
                    s256, s128 = s128 & v, s128 ^ v
                else:
                    u64a = u
            else:
                t32a = t
        else:
            o_prev = o
    return [np.uint32(ptp_planes[0]), ones, twos, fours, eights, s16, s32, s64, s128, s256]


def _bitslice(T):
    """32 pooled rows [32, 256] bits -> 267 lane words (bit rows, zero row) and the nine planes of 256 - popc."""
    rows = np.zeros(267, np.uint32)
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    rows[:256] = (T.astype(np.uint64) * w[:, None]).sum(0).astype(np.uint32)
    ptp = 256 - T.sum(1)
    planes = [np.uint32((((ptp >> p) & 1).astype(np.uint64) * w).sum()) for p in range(9)]
    return rows, planes


@pytest.mark.parametrize("gran", [0, 1, 2])
def test_list_walk_equals_popcount_distance(gran):
    rng = np.random.default_rng(gran)
    for trial in range(120):
        T = rng.integers(0, 2, (32, 256), dtype=np.uint8)
        if trial % 3 == 0:
            T[:, :200] = 0
        q = rng.integers(0, 2, 256, dtype=np.uint8)
        if trial % 2:
            q[rng.integers(0, 256, int(rng.integers(0, 250)))] = trial % 4 == 1
        rows, planes = _bitslice(T)
        A = int(q.sum())
        inv = A > 128
        bits = [b for b in range(256) if q[b] != inv]
        n = len(bits)
        assert n <= 128
        lst = bits + [ZERO_ROW] * (128 - n)
        nh = (n + 7) // 8 if gran == 2 else 2 * ((n + 15) // 16)
        P = _walk(rows, planes, lst, nh, gran)
        s = sum(((np.array([int(p) for p in P])[:, None] >> np.arange(32)) & 1)[p] << p for p in range(10))
        d = A - 256 + s if inv else A + 256 - s
        assert np.array_equal(d, (T ^ q).sum(1)), (trial, n, nh)


def test_majority_flip_keeps_distances_and_shortens_lists():
    rng = np.random.default_rng(5)
    p_bit = rng.choice([0.05, 0.2, 0.5, 0.8, 0.95], 256)
    pool = rng.random((3000, 256)) < p_bit
    q = rng.random((200, 256)) < p_bit
    n_s = min(len(pool), 8192)
    maj = 2 * pool[(np.arange(n_s) * len(pool)) // n_s].sum(0) > n_s      # knn5_flip_kernel
    d0 = (q[:, None, :] ^ pool[None, :, :]).sum(2)
    d1 = ((q ^ maj)[:, None, :] ^ (pool ^ maj)[None, :, :]).sum(2)
    assert np.array_equal(d0, d1)
    a0, a1 = q.sum(1), (q ^ maj).sum(1)
    assert np.minimum(a1, 256 - a1).mean() < 0.8 * np.minimum(a0, 256 - a0).mean()


def test_dealing_is_a_bijection_and_balances_the_warps():
    rng = np.random.default_rng(6)
    for _ in range(50):
        nh = rng.integers(0, 17, 128)
        nh[rng.integers(0, 128, 10)] = 0                                  # rows beyond the launch's last query
        # rank = number of rows with a longer list, or an equal one and a smaller index (knn5_kernel, the dealing step)
        rank = np.array([int(((nh > nh[t]) | ((nh == nh[t]) & (np.arange(128) < t))).sum()) for t in range(128)])
        assert sorted(rank) == list(range(128))
        rnd, j = rank // 16, rank % 16
        slot = np.where(rnd & 1, 15 - j, j) * 8 + rnd
        assert sorted(slot) == list(range(128))
        per_warp = np.zeros(16)
        np.add.at(per_warp, slot // 8, nh)
        natural = nh.reshape(16, 8).sum(1)
        assert per_warp.max() - per_warp.min() <= 16                      # within one query's walk
        assert per_warp.max() <= natural.max()


@pytest.mark.parametrize("nq", [1, 127, 128, 129, 130, 257, 18944, 18945, 300000])
def test_tile_map_covers_the_query_range_once(nq):
    n_tiles = (nq + 127) // 128
    ql, tile = np.meshgrid(np.arange(128), np.arange(n_tiles), indexing="ij")
    row = (ql * n_tiles + tile).ravel()
    row = row[row < nq]
    assert len(row) == nq and np.array_equal(np.sort(row), np.arange(nq))
    # the merge kernel's inverse:  ql = row / n_tiles,  tile = row % n_tiles
    assert np.array_equal(row // n_tiles * n_tiles + row % n_tiles, row)
