"""Round-2 GPU parity tests of the device-driven frame stream and of K8 v5 (bit-sliced Hamming k-NN), all through the C ABI.

* K8 v5 (default) against K8 v4 (cfg.knn_impl = 4, the XOR/POPC kernel of round 1, itself bit-exact vs the oracle in
  test_gpu_knn.py) on shapes the CPU oracle cannot finish in seconds, and against the oracle on shapes that exercise v5's own edge
  cases (pool sizes around the 4096-row slab, queries with > 128 set bits, all-zero / all-one descriptors, split tiles, every
  list length on a biased pool);
* submit / collect == match_frames, several tickets in flight, tickets collected out of order, blank frames (zero keypoints)
  inside a stream, a stream long enough to cross an epoch (2048 frames) on small frames;
* the changed-frame chain across calls of growing size (slot 0 must survive a regrowth);
* pool replication between two ctxs on one GPU (export/import and reserve -> device view -> commit), ORB256 and SIFT128.
"""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


# ---- K8 v5 -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt,k", [
    (1, 4095, 30), (2, 4096, 30), (3, 4097, 30), (129, 8192, 30), (127, 12289, 30), (1000, 50000, 30), (40, 300000, 30),
    (5000, 9000, 5), (777, 4096 * 3 + 1, 32),
])
def test_v5_matches_oracle_around_slab_sizes(nq, nt, k):
    import slideo_b200
    pool = synth.hamming_pool(nt, seed=21, dup_frac=0.02)
    q = synth.hamming_queries(pool, nq, seed=22)
    with slideo_b200.Context() as c:
        gi, gd = c.bf_knn_hamming(q, pool, k)
    oi, od = oracle.bf_knn_hamming(q, pool, k)
    assert np.array_equal(gd, od) and np.array_equal(gi, oi)


def test_v5_dense_sparse_and_constant_descriptors():
    """Popcounts 0, 256, < 128, > 128 (complemented lists), both on the query and on the pool side."""
    import slideo_b200
    rng = np.random.default_rng(4)
    def with_density(n, p):
        return np.packbits(rng.random((n, 256)) < p, axis=1)
    pool = np.concatenate([with_density(3000, 0.5), with_density(700, 0.05), with_density(700, 0.95), np.zeros((40, 32), np.uint8),
                           np.full((40, 32), 255, np.uint8), with_density(1000, 0.5)])
    q = np.concatenate([with_density(50, 0.5), with_density(50, 0.02), with_density(50, 0.98), np.zeros((3, 32), np.uint8),
                        np.full((3, 32), 255, np.uint8), with_density(20, 129 / 256), with_density(20, 127 / 256), pool[::97]])
    with slideo_b200.Context() as c:
        for k in (30, 1, 32):
            gi, gd = c.bf_knn_hamming(q, pool, k)
            oi, od = oracle.bf_knn_hamming(q, pool, k)
            assert np.array_equal(gd, od) and np.array_equal(gi, oi), k


def test_v5_every_list_length_on_a_biased_pool():
    """K8 v5 XORs pool and queries with the pool's majority vector and walks a query's list only as far as it is long: one query for
    every list length 0..128 on either side of the complement switch (257 queries = three partly filled tiles whose rows are dealt
    to the warps by length), on a pool whose bits are biased like descriptors of slides."""
    import slideo_b200
    rng = np.random.default_rng(11)
    p_bit = rng.choice([0.03, 0.2, 0.5, 0.8, 0.97], 256)
    pool_bits = rng.random((6000, 256)) < p_bit                     # <= 8192 rows: the kernel's sample is the whole pool
    maj = 2 * pool_bits.sum(0) > len(pool_bits)
    pool = np.packbits(pool_bits, axis=1)
    q_bits = np.zeros((257, 256), bool)
    for n in range(257):                                            # popcount n after the flip: list length min(n, 256 - n)
        q_bits[n, rng.permutation(256)[:n]] = True
    q = np.packbits(q_bits ^ maj, axis=1)
    q = np.concatenate([q, pool[::501]])                            # and some exact hits (distance 0)
    with slideo_b200.Context() as c:
        for k in (30, 7):
            gi, gd = c.bf_knn_hamming(q, pool, k)
            oi, od = oracle.bf_knn_hamming(q, pool, k)
            assert np.array_equal(gd, od) and np.array_equal(gi, oi), k


def test_v5_equals_v4_on_large_shapes():
    import slideo_b200
    import torch
    k = 30
    for nq, nt in ((40000, 120000), (300, 1_000_000), (19000, 4096 * 5)):
        pool = synth.hamming_pool(nt, seed=31, dup_frac=0.01)
        q = synth.hamming_queries(pool, nq, seed=32)
        dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(pool).cuda()
        keys = []
        for impl in (0, 4):
            out = torch.empty((nq, k), dtype=torch.int32, device="cuda")
            with slideo_b200.Context(slideo_b200.default_config(knn_impl=impl)) as c:
                c.bf_knn_hamming_device(dq.data_ptr(), nq, dt.data_ptr(), nt, k, out.data_ptr())
                c.synchronize()
            keys.append(out.cpu().numpy())
        assert np.array_equal(keys[0], keys[1]), (nq, nt)


# ---- the device-driven stream ----------------------------------------------------------------------------------------------------
NPAGES = 5


@pytest.fixture(scope="module")
def scene():
    pages = [synth.make_page(p) for p in range(NPAGES)]
    frames = np.stack([synth.make_frame(f, NPAGES, pages) for f in range(12)])
    frames[4] = 200                                    # a blank frame: zero keypoints inside the stream
    frames[7, :, :, :] = frames[7, :1, :1, :]          # another one
    page_desc = [oracle.orb_detect_and_compute(p)[2] for p in pages]
    expect = []
    for f in frames:
        d = oracle.orb_detect_and_compute(oracle.gray_from_bgr(f))[2]
        best, votes, _ = oracle.match_frame(d, page_desc)
        expect.append((best, votes, len(d)))
    return pages, frames, np.array(expect, np.int32)


def test_submit_collect_equals_oracle_and_sync_path(scene):
    import slideo_b200
    pages, frames, expect = scene
    assert expect[4, 2] == 0 and expect[7, 2] == 0 and expect[4, 0] == -1
    with slideo_b200.Context(slideo_b200.default_config(max_batch=4)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        pin = slideo_b200.PinnedBuffer(frames.nbytes)
        pin.array[:] = frames.reshape(-1)
        fb = frames[0].nbytes
        # three tickets in flight over one continuous query stream, collected out of order
        t1 = c.submit_frames_bgr8_ptr(pin.ptr, 5, 1920, 1080)
        t2 = c.submit_frames_bgr8_ptr(pin.ptr + 5 * fb, 1, 1920, 1080)
        t3 = c.submit_frames_bgr8_ptr(pin.ptr + 6 * fb, 6, 1920, 1080)
        r2 = c.collect(t2, 1)
        r1 = c.collect(t1, 5)
        r3 = c.collect(t3, 6)
        got = np.concatenate([r1, r2, r3])
        assert np.array_equal(got, expect)
        assert np.array_equal(c.match_frames_bgr8(frames), expect)
        # a collected ticket is gone; an unknown one is an argument error
        with pytest.raises(slideo_b200.SlideoError) as e:
            c.collect(t1, 5)
        assert e.value.status == slideo_b200.ffi.E_INVALID_ARG
        # device-resident submit, repeated: the stream keeps going across calls
        import torch
        d = torch.from_numpy(frames).cuda()
        ts = [c.submit_frames_bgr8_device(d.data_ptr(), len(frames), 1920, 1080) for _ in range(3)]
        for t in ts:
            assert np.array_equal(c.collect(t, len(frames)), expect)
        tm = c.timings()
        assert tm["frames"] == 5 * len(frames) and tm["knn_pairs"] == int(expect[:, 2].sum()) * 5 * c.pool_info()[0]
        pin.close()


def test_stream_crosses_an_epoch_on_small_frames():
    """2048 frames per epoch: 2500 small frames in one call and as 5 tickets must give the same rows; sampled rows == oracle."""
    import slideo_b200
    rng = np.random.default_rng(8)
    def tex(seed, h, w):
        r = np.random.default_rng(seed)
        small = r.integers(0, 256, (h // 8 + 1, w // 8 + 1)).astype(np.float32)
        big = np.kron(small, np.ones((8, 8), np.float32))[:h, :w]
        big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, (1, 1), (0, 1))) / 4
        return big.astype(np.uint8)
    h, w, n = 240, 320, 2500
    pages = [tex(i, h, w) for i in range(3)]
    base = np.stack([np.stack([pages[i % 3]] * 3, axis=2) for i in range(8)])
    frames = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        frames[i] = np.clip(base[i % 8].astype(np.int16) + rng.integers(-2, 3, (h, w, 1)), 0, 255)
    with slideo_b200.Context(slideo_b200.default_config(max_batch=64)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        whole = c.match_frames_bgr8(frames)
        tickets = [c.submit_frames_bgr8(frames[i:i + 500]) for i in range(0, n, 500)]
        parts = np.concatenate([c.collect(t, 500) for t in tickets])
    assert np.array_equal(whole, parts)
    page_desc = [oracle.orb_detect_and_compute(p)[2] for p in pages]
    for i in (0, 1, 2047, 2048, 2049, n - 1):
        d = oracle.orb_detect_and_compute(oracle.gray_from_bgr(frames[i]))[2]
        best, votes, _ = oracle.match_frame(d, page_desc)
        assert tuple(whole[i]) == (best, votes, len(d)), i


# ---- changed-frame chain across calls of growing size -------------------------------------------------------------
def test_prefilter_chain_survives_growing_calls():
    import slideo_b200
    pages = [synth.make_page(p) for p in range(2)]
    frames = np.stack([synth.make_frame(f, 2, pages) for f in (0, 0, 1, 1, 0)])
    frames[1] = np.clip(frames[0].astype(np.int16) + 1, 0, 255)
    with slideo_b200.Context(slideo_b200.default_config(max_batch=8)) as c:
        ch_all, sim_all = c.mark_changed_bgr8(frames, reset=True)
    with slideo_b200.Context(slideo_b200.default_config(max_batch=8)) as c:
        ch1, sim1 = c.mark_changed_bgr8(frames[:1], reset=True)
        ch2, sim2 = c.mark_changed_bgr8(frames[1:])
        # a page added in between must not disturb the chain either
    assert np.array_equal(np.concatenate([sim1, sim2]).view(np.uint32), sim_all.view(np.uint32))
    assert np.array_equal(np.concatenate([ch1, ch2]), ch_all)
    assert not ch_all[1] and ch_all[2]


# ---- pool replication between two ctxs on one GPU -----------------------------------------------------
def _replicate_device(a, b):
    """pool_reserve -> device views -> device-to-device copies -> (points) -> commit: what sharding.broadcast_pool_device does over
    NCCL, here with plain copies between two ctxs of one GPU."""
    import torch
    from slideo_b200.sharding import _DevView
    dev = torch.device("cuda", 0)
    n, p = a.pool_info()
    b.pool_reserve(n, p)
    sd, sbytes, so, obytes = a.pool_device_view()
    dd, dbytes, do, dobytes = b.pool_device_view()
    assert (sbytes, obytes) == (dbytes, dobytes)
    torch.as_tensor(_DevView(do, obytes), device=dev).copy_(torch.as_tensor(_DevView(so, obytes), device=dev))
    if sbytes:
        torch.as_tensor(_DevView(dd, sbytes), device=dev).copy_(torch.as_tensor(_DevView(sd, sbytes), device=dev))
    sp, pbytes, has = a.pool_points_device_view()
    if has and pbytes:
        dp, dpbytes, _ = b.pool_points_device_view()
        assert dpbytes == pbytes
        torch.as_tensor(_DevView(dp, pbytes), device=dev).copy_(torch.as_tensor(_DevView(sp, pbytes), device=dev))
        b.pool_points_device_view(received=True)
    ss, sm_bytes, pw, ph = a.pool_pages_device_view()
    if pw > 0 and sm_bytes:
        ds, dsm_bytes, _, _ = b.pool_pages_device_view(pw, ph)
        assert dsm_bytes == sm_bytes
        torch.as_tensor(_DevView(ds, sm_bytes), device=dev).copy_(torch.as_tensor(_DevView(ss, sm_bytes), device=dev))
    torch.cuda.synchronize()
    b.pool_commit()


@pytest.mark.parametrize("kind", ["orb", "sift"])
def test_pool_replication_gives_identical_results(scene, kind):
    import slideo_b200
    pages, frames, expect = scene
    if kind == "sift":
        pages = [np.ascontiguousarray(p[:400, :600]) for p in pages[:3]]
        frames = np.ascontiguousarray(frames[:4, :400, :600])
        cfg = dict(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=2)
    else:
        cfg = dict(max_batch=4, geometric_verification=2)
    with slideo_b200.Context(slideo_b200.default_config(**cfg)) as a, slideo_b200.Context(slideo_b200.default_config(**cfg)) as b, \
            slideo_b200.Context(slideo_b200.default_config(**{**cfg, "geometric_verification": 0})) as c:
        for p in pages:
            a.add_page_gray8(p)
        a.finalize_pool()
        ra = a.match_frames_bgr8(frames)
        if kind == "orb":
            assert np.array_equal(ra, expect)
        _replicate_device(a, b)
        assert np.array_equal(b.match_frames_bgr8(frames), ra)
        if kind == "orb":
            assert a.get_verification(0, len(frames)) == b.get_verification(0, len(frames))
            da, db = a.get_decisions(0, len(frames)), b.get_decisions(0, len(frames))
            assert [d["image"] for d in da] == [d["image"] for d in db] and any(d["image"] >= 0 for d in da)
            assert repr(da) == repr(db)            # rated pages, similarities and refined matrices bit for bit
        desc, offs = a.pool_export()
        c.pool_import(desc, offs)
        assert c.pool_info() == a.pool_info()
        assert np.array_equal(c.match_frames_bgr8(frames), ra)


def test_stream_geometry_change_progress_callback_and_dropped_ticket(scene):
    """Two geometries in flight at once (the stream closes its epoch and opens a new one), a ticket collected into NULL, and the
    progress callback (crates/matching/src/progress.rs:3-17 as a C function pointer)."""
    import ctypes
    import slideo_b200
    from slideo_b200 import ffi
    pages, frames, expect = scene
    small_pages = [np.ascontiguousarray(p[200:680, 300:940]) for p in pages]
    small = np.ascontiguousarray(frames[:6, 200:680, 300:940])
    with slideo_b200.Context(slideo_b200.default_config(max_batch=4)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        want_small = c.match_frames_bgr8(small)
        seen = []
        c.set_progress_callback(lambda a, b, m: seen.append((a, b, m)))
        big = np.ascontiguousarray(frames)
        t1 = c.submit_frames_bgr8(big)
        t2 = c.submit_frames_bgr8(small)
        t3 = c.submit_frames_bgr8(big)
        assert np.array_equal(c.collect(t1, len(big)), expect)
        assert seen and all(b == len(big) and a <= b for a, b, _ in seen) and seen[-1][2].startswith("Processing")
        assert seen[-1][0] == len(big)
        # drop ticket 2 (out = NULL), then ticket 3 must still be right
        got = ctypes.c_int32()
        c._ck(c._lib.slideo_b200_collect(c._h, t2, None, 0, ctypes.byref(got)))
        assert got.value == len(small)
        c.set_progress_callback(None)
        n_seen = len(seen)
        assert np.array_equal(c.collect(t3, len(big)), expect)
        assert len(seen) == n_seen
        assert np.array_equal(c.match_frames_bgr8(small), want_small)
    # the small crops equal the oracle too (votes may be low: the check is parity, not recognition)
    page_desc = [oracle.orb_detect_and_compute(p)[2] for p in pages]
    for i in (0, 5):
        d = oracle.orb_detect_and_compute(oracle.gray_from_bgr(small[i]))[2]
        best, votes, _ = oracle.match_frame(d, page_desc)
        assert tuple(want_small[i]) == (best, votes, len(d))
    del small_pages
