"""K11 parity: the GPU SIFT (slideo_b200/csrc/sift.cu, through the C ABI) vs the oracle restatement (oracle/sift_oracle.c, pinned
against cv2.SIFT_create().detectAndCompute in tests/test_oracle_sift.py).

Everything is compared bit for bit: every Gaussian layer of the scale space, the keypoint count, coordinates, packed octaves,
responses, angles, sizes and all 128 descriptor elements.  The three libm values of the algorithm (exp2f for the size, cosf /
sinf for the descriptor rotation) are evaluated on the device with glibc's own algorithms (sift.cu: glibc_exp2f / glibc_sincosf),
so nothing is left to a tolerance; frame-level results (best slide AND vote count) equal the oracle pipeline exactly.  The only
tolerance left in this file is cv2's tolerance against ITSELF in test_gpu_sift_equals_cv2_directly."""
import numpy as np
import pytest

import oracle
import slideo_b200
import synth

pytestmark = pytest.mark.gpu


def textured(seed, h, w, cell=8):
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (h // cell + 2, w // cell + 2)).astype(np.float32)
    big = np.kron(small, np.ones((cell, cell), np.float32))[:h, :w]
    for _ in range(3):
        big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, (1, 1), (0, 1))) / 4
    return np.clip(big, 0, 255).astype(np.uint8)


def check_against_oracle(got, ref):
    kf, oc, de = got
    rkf, roc, rde = ref
    assert len(kf) == len(rkf), (len(kf), len(rkf))
    assert np.array_equal(oc, roc), "packed octaves differ"
    assert np.array_equal(kf[:, :2].view(np.uint32), rkf[:, :2].view(np.uint32)), "keypoint coordinates differ"
    assert np.array_equal(kf[:, 4].view(np.uint32), rkf[:, 4].view(np.uint32)), "responses differ"
    assert np.array_equal(kf[:, 3].view(np.uint32), rkf[:, 3].view(np.uint32)), "angles differ"
    assert np.array_equal(kf[:, 2].view(np.uint32), rkf[:, 2].view(np.uint32)), "sizes differ (exp2f)"
    assert np.array_equal(de.view(np.uint32), rde.view(np.uint32)), "descriptors differ (cosf / sinf of the orientation)"
    return len(kf)


@pytest.mark.parametrize("shape,seed", [((250, 333), 1), ((120, 200), 2), ((67, 121), 3), ((400, 640), 4)])
def test_scale_space_is_bit_exact(ctx, shape, seed):
    """Every Gaussian layer of every octave (upsample, blur chain incl. OpenCV's unfused remainder columns, INTER_NEAREST half)."""
    g = textured(seed, *shape)
    ctx.extract_sift(g)
    n_oct = oracle.lib().sift_num_octaves(shape[1], shape[0])
    for o in range(n_oct):
        for layer in ((0, 1, 2, 3, 4, 5) if o == 0 else (0, 3, 5)):
            mine = ctx.debug_fetch_sift(o, layer)
            ref = oracle.sift_pyramid_image(g, 0, o, layer)
            assert mine.shape == ref.shape
            assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), f"octave {o} layer {layer}"


@pytest.mark.parametrize("shape,seed", [((250, 333), 1), ((120, 200), 2), ((67, 121), 3), ((400, 640), 4), ((64, 64), 5)])
def test_detect_and_compute_equals_oracle(ctx, shape, seed):
    g = textured(seed, *shape)
    n = check_against_oracle(ctx.extract_sift(g), oracle.sift_detect_and_compute(g))
    assert n > 20 or shape == (64, 64)


def test_bgr_input_and_flat_image(ctx):
    g = textured(7, 200, 300)
    rng = np.random.default_rng(7)
    bgr = np.stack([g, np.roll(g, 3, 1), rng.integers(0, 256, g.shape, dtype=np.uint8)], axis=2)
    check_against_oracle(ctx.extract_sift(bgr), oracle.sift_detect_and_compute(oracle.gray_from_bgr(bgr)))
    flat = np.full((100, 150), 200, np.uint8)
    kf, oc, de = ctx.extract_sift(flat)
    assert len(kf) == 0 and de.shape == (0, 128)


def test_synthetic_page_and_frame_full_size(ctx):
    """The BASELINE geometry: a 2001 x 1125 page and a 1080p BGR frame (about 6 k keypoints each)."""
    page = synth.make_page(3)
    n = check_against_oracle(ctx.extract_sift(page), oracle.sift_detect_and_compute(page))
    assert n > 2000
    frame = synth.make_frame(3, 50)
    n = check_against_oracle(ctx.extract_sift(frame), oracle.sift_detect_and_compute(oracle.gray_from_bgr(frame)))
    assert n > 2000


def test_sift_frame_path_equals_oracle_pipeline():
    """SIFT128 ctx: pages as images -> K11 -> pool; frames -> K11 -> K10 -> vote.  (best_slide, votes, n_keypoints) equal the
    oracle's SIFT + BFMatcher(NORM_L2) + vote on the same images."""
    pages = [np.ascontiguousarray(synth.make_page(p)[60:700, 100:1060]) for p in range(4)]        # 960 x 640 crops
    rng = np.random.default_rng(11)
    frames = []
    for p in (2, 0, 3, 1, 2):
        g = pages[p].astype(np.int16) + rng.integers(-4, 5, pages[p].shape)
        g = np.clip(g, 0, 255).astype(np.uint8)
        frames.append(np.stack([g, g, g], axis=2))
    frames = np.stack(frames)
    with slideo_b200.Context(slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=2, keep_matches=1)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        pool, offs = c.pool_export()
        res = c.match_frames_bgr8(frames)
        t = c.timings()
        rows0 = c.get_matches(0)
    ref_pages = [oracle.sift_detect_and_compute(p)[2] for p in pages]
    ref_pool = np.concatenate(ref_pages)
    assert list(offs) == list(np.concatenate([[0], np.cumsum([len(d) for d in ref_pages])]))
    assert np.array_equal(pool.reshape(-1, 128).view(np.uint32), ref_pool.view(np.uint32))
    for i, f in enumerate(frames):
        kf, oc, de = oracle.sift_detect_and_compute(oracle.gray_from_bgr(f))
        idx, dist = oracle.bf_knn_l2(de, ref_pool, 30)
        best, votes, _ = oracle.vote(idx, dist, offs)
        assert res[i, 2] == len(kf)
        assert res[i, 0] == best == (2, 0, 3, 1, 2)[i]
        assert int(res[i, 1]) == votes, (res[i], votes)
    assert t["kernel_launches"] > 0 and t["knn_launches"] > 0
    assert rows0.shape == (res[0, 2], 30)


def test_sift_batched_equals_single(ctx):
    """Batch composition must not change the result: frames through the batched path vs one by one (bitwise)."""
    pages = [textured(20 + p, 300, 400) for p in range(3)]
    rng = np.random.default_rng(3)   # noise: an exact copy has best distance 0 and, by the reference's rule (lib.rs:275), casts no vote
    noisy = [np.clip(pages[p].astype(np.int16) + rng.integers(-3, 4, pages[p].shape), 0, 255).astype(np.uint8) for p in (1, 2, 0, 1, 0)]
    frames = np.stack([np.stack([g] * 3, axis=2) for g in noisy])
    out = []
    for mb in (1, 4):
        with slideo_b200.Context(slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=mb)) as c:
            for p in pages:
                c.add_page_gray8(p)
            c.finalize_pool()
            out.append((c.pool_export()[0].copy(), c.match_frames_bgr8(frames).copy()))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert list(out[0][1][:, 0]) == [1, 2, 0, 1, 0]


def test_sift_edge_cases_flat_frames_and_empty_pages():
    """A flat page adds no descriptors (empty page in the pool), a flat frame has no keypoints (best_slide -1, 0 votes), and a
    batch may mix both with ordinary frames; k = 7 exercises k < 30 on the SIFT path."""
    pages = [textured(40, 200, 260), np.full((200, 260), 255, np.uint8), textured(41, 200, 260)]
    rng = np.random.default_rng(9)
    noisy = lambda p: np.clip(pages[p].astype(np.int16) + rng.integers(-3, 4, pages[p].shape), 0, 255).astype(np.uint8)
    grays = [noisy(2), np.full((200, 260), 17, np.uint8), noisy(0), np.full((200, 260), 200, np.uint8), noisy(2)]
    frames = np.stack([np.stack([g] * 3, axis=2) for g in grays])
    with slideo_b200.Context(slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=3, knn_k=7)) as c:
        counts = [c.add_page_gray8(p) for p in pages]
        c.finalize_pool()
        pool, offs = c.pool_export()
        res = c.match_frames_bgr8(frames)
        assert c.match_frames_bgr8(frames[:0]).shape == (0, 3)
    assert counts[1] == 0 and offs[1] == offs[2]
    ref_pool = np.concatenate([oracle.sift_detect_and_compute(p)[2] for p in pages])
    for i, g in enumerate(grays):
        de = oracle.sift_detect_and_compute(g)[2]
        if len(de) == 0:
            assert tuple(res[i]) == (-1, 0, 0)
            continue
        idx, dist = oracle.bf_knn_l2(de, ref_pool, 7)
        best, votes, _ = oracle.vote(idx, dist, offs)
        assert res[i, 0] == best == (2, -1, 0, -1, 2)[i] and res[i, 2] == len(de)
        assert int(res[i, 1]) == votes


def test_gpu_sift_equals_cv2_directly(ctx):
    """The GPU against the real OpenCV (cv2.SIFT_create().detectAndCompute, IPP off, one thread -- OpenCV's own deterministic code
    path), at the tolerance cv2 shows against itself (tests/test_oracle_sift.py): identical keypoint count, coordinates and packed
    octaves; size / angle / response bit-identical on >= 95 % and within 1e-5 relative; descriptors +-1 on < 0.1 % of the elements."""
    cv2 = pytest.importorskip("cv2")
    ipp, thr = cv2.ipp.useIPP(), cv2.getNumThreads()
    cv2.ipp.setUseIPP(False)
    cv2.setNumThreads(1)
    try:
        for img in (textured(51, 203, 271), np.ascontiguousarray(synth.make_page(5)[100:500, 40:680])):
            kp, d = cv2.SIFT_create().detectAndCompute(img, None)
            gkf = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kp], np.float32).reshape(-1, 5)
            goc = np.array([k.octave for k in kp], np.int32)
            kf, oc, de = ctx.extract_sift(img)
            assert len(kf) == len(gkf) > 100
            assert np.array_equal(oc, goc)
            assert np.array_equal(kf[:, :2], gkf[:, :2])
            for col in (2, 3, 4):
                assert np.mean(kf[:, col] == gkf[:, col]) >= 0.95
                assert np.allclose(kf[:, col], gkf[:, col], rtol=1e-5, atol=1e-4 if col == 3 else 0)
            diff = np.abs(de.astype(np.int32) - d.astype(np.int32))
            assert diff.max() <= 1 and np.mean(diff == 0) >= 0.999
    finally:
        cv2.ipp.setUseIPP(ipp)
        cv2.setNumThreads(thr)


@pytest.mark.parametrize("shape", [(9, 8), (16, 33), (33, 47)])
def test_tiny_images_equal_oracle(ctx, shape):
    """Images smaller than the blur kernels (every tile is a border tile, reflect-101 wraps several times) and with 2-3 octaves only."""
    g = textured(60 + shape[0], shape[0], shape[1], cell=4)
    check_against_oracle(ctx.extract_sift(g), oracle.sift_detect_and_compute(g))
    n_oct = oracle.lib().sift_num_octaves(shape[1], shape[0])
    mine, ref = ctx.debug_fetch_sift(n_oct - 1, 5), oracle.sift_pyramid_image(g, 0, n_oct - 1, 5)
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
