"""K12 parity: the RANSAC gate (lib.rs:284-333) on the GPU == the oracle pipeline (exact k-NN rows -> 1.05 votes grouped by
slide in (query, rank) order -> top-40 -> oracle/ransac_oracle.c, itself pinned against cv2.estimateAffinePartial2D ->
sort / truncate(10) / retain gates) on the same seeded pages and frames."""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu

NPAGES, NFRAMES = 7, 8


@pytest.fixture(scope="module")
def scene():
    pages = [synth.make_page(p) for p in range(NPAGES)]
    frames = np.stack([synth.make_frame(f, NPAGES, pages) for f in range(NFRAMES)])
    feats = [oracle.orb_detect_and_compute(p) for p in pages]
    page_desc = [f[2] for f in feats]
    pool_pts = np.concatenate([f[1][:, :2] for f in feats])
    offs = np.zeros(NPAGES + 1, np.int32)
    offs[1:] = np.cumsum([len(d) for d in page_desc])
    pool = np.concatenate(page_desc)
    expect = []
    for f in frames:
        ki, kf, d = oracle.orb_detect_and_compute(oracle.gray_from_bgr(f))
        idx, dist = oracle.bf_knn_hamming(d, pool, 30)
        expect.append(oracle.verify_frame(idx, dist.astype(np.float32), offs, kf[:, :2], pool_pts))
    return pages, frames, feats, expect


@pytest.mark.parametrize("via_features", [False, True])
def test_verification_equals_oracle(scene, via_features):
    import slideo_b200
    pages, frames, feats, expect = scene
    with slideo_b200.Context(slideo_b200.default_config(geometric_verification=1, max_batch=3)) as c:
        for p, (ki, kf, d) in zip(pages, feats):
            if via_features:
                c.add_page_features(d, kf[:, :2])
            else:
                c.add_page_gray8(p)
        c.finalize_pool()
        res = c.match_frames_bgr8(frames)
        got = c.get_verification(0, NFRAMES)
        t = c.timings()
    assert t["ms_verify"] > 0
    for f in range(NFRAMES):
        assert got[f]["cand"] == [tuple(int(v) for v in x) for x in expect[f]["cand"]], f"frame {f} candidates"
        assert got[f]["survivors"] == [tuple(int(v) for v in x) for x in expect[f]["survivors"]], f"frame {f} survivors"
        # the head of the vote ranking is the hot path's (best_slide, votes)
        if got[f]["cand"]:
            assert (res[f, 0], res[f, 1]) == got[f]["cand"][0][:2]
        truth = synth.frame_truth(f, NPAGES)
        if truth >= 0:
            assert got[f]["survivors"] and got[f]["survivors"][0][0] == truth     # the shown page passes the gate, first
        else:
            assert not got[f]["survivors"]                                         # clutter frames: nothing passes


def test_verification_needs_points():
    import slideo_b200
    with slideo_b200.Context(slideo_b200.default_config(geometric_verification=1)) as c:
        c.add_page_descriptors(synth.hamming_pool(100, seed=3))
        c.finalize_pool()
        with pytest.raises(slideo_b200.SlideoError) as e:
            c.match_frames_bgr8(synth.make_frame(0, 1)[None])
        assert e.value.status == slideo_b200.ffi.E_STATE


def test_matcher_mirror_uses_the_gate(scene):
    """create_video_matcher -> match_images_with_video -> process with the RANSAC gate deciding (lib.rs:329-333)."""
    import slideo_b200
    pages, frames, _, expect = scene
    vm = slideo_b200.B200ImageVideoMatcher().create_video_matcher(pages)
    task = vm.match_images_with_video([(frames[i], 5.0 * i, 125 * i) for i in range(NFRAMES)])
    task.prefilter = False
    out = task.process()
    want, last = [], "start"
    for i in range(NFRAMES):
        img = expect[i]["survivors"][0][0] if expect[i]["survivors"] else None
        if last != "start" and last == img:
            continue
        last = img
        want.append((125 * i, img))
    got = [(m.video_frame_idx, None if m.image is None else next(j for j, p in enumerate(pages) if p is m.image)) for m in out]
    assert got == want
    vm.ctx.close()


def test_photometric_gate_equals_oracle(scene):
    """cfg.geometric_verification == 2: LM-refined matrix -> warpAffine(nearest, inverse) -> INTER_AREA -> similarity gate
    (lib.rs:335-389) against oracle/photometric.py (pinned against cv2.warpAffine / cv2.estimateAffinePartial2D)."""
    import slideo_b200
    from oracle import photometric as ph
    pages, frames, feats, expect = scene
    page_desc = [f[2] for f in feats]
    pool_pts = np.concatenate([f[1][:, :2] for f in feats])
    offs = np.zeros(NPAGES + 1, np.int32)
    offs[1:] = np.cumsum([len(d) for d in page_desc])
    pool = np.concatenate(page_desc)
    with slideo_b200.Context(slideo_b200.default_config(geometric_verification=2, max_batch=3)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        res = c.match_frames_bgr8(frames)
        ver = c.get_verification(0, NFRAMES)
        dec = c.get_decisions(0, NFRAMES)
        import torch
        d = torch.from_numpy(frames).cuda()
        res_d = c.match_frames_bgr8_device(d.data_ptr(), NFRAMES, 1920, 1080)
        dec_d = c.get_decisions(0, NFRAMES)
    assert np.array_equal(res, res_d)
    for f in range(NFRAMES):
        ki, kf, dd = oracle.orb_detect_and_compute(oracle.gray_from_bgr(frames[f]))
        idx, dist = oracle.bf_knn_hamming(dd, pool, 30)
        want = ph.decide_frame(idx, dist.astype(np.float32), offs, kf[:, :2], pool_pts, frames[f], pages)
        assert ver[f]["survivors"] == [tuple(int(v) for v in x) for x in want["survivors"]]
        assert dec[f]["image"] == want["image"] == dec_d[f]["image"], f"frame {f}"
        assert [p for p, _ in dec[f]["rated"]] == [p for p, _ in want["rated"]]
        # similarities: bit-identical whenever the refined matrix lands on the same fixed-point coordinates (it does unless a
        # coordinate sits within ~1e-9 of a rounding boundary); the matrix itself agrees to 1e-9 (different 4x4 solvers)
        for (p, s), (_, ws) in zip(dec[f]["rated"], want["rated"]):
            assert abs(float(s) - float(ws)) <= 1e-6
        assert [float(s) for _, s in dec[f]["rated"]] == [float(s) for _, s in dec_d[f]["rated"]]
        for j, (p, _) in enumerate(want["survivors"]):
            m = want["matrices"][p]
            assert np.allclose(dec[f]["refined"][j], [m[0, 0], m[1, 0], m[0, 2], m[1, 2]], rtol=0, atol=1e-9)
        truth = synth.frame_truth(f, NPAGES)
        assert dec[f]["image"] == (truth if truth >= 0 else -1)


def test_photometric_gate_with_mixed_page_sizes():
    """A deck whose pages differ in size: the reference warps the frame to each slide's own size and compares small images of that
    slide's geometry (lib.rs:339-351, slide_info.img.size()); the library keeps one geometry class per page size.  Decisions,
    survivors and similarities against the oracle chain on the same pages; replication (one page size per deck) reports 0 x 0."""
    import cv2
    import slideo_b200
    from oracle import photometric as ph
    npages = 4
    base = [synth.make_page(p) for p in range(npages)]
    sizes = [None, (1600, 900), None, (1440, 810)]
    pages = [b if s is None else cv2.resize(b, s, interpolation=cv2.INTER_AREA) for b, s in zip(base, sizes)]
    frames = np.stack([synth.make_frame(f, npages, base) for f in range(6)])   # frames show the full-size renderings
    feats = [oracle.orb_detect_and_compute(p) for p in pages]
    pool = np.concatenate([f[2] for f in feats])
    pool_pts = np.concatenate([f[1][:, :2] for f in feats])
    offs = np.zeros(npages + 1, np.int32)
    offs[1:] = np.cumsum([len(f[2]) for f in feats])
    with slideo_b200.Context(slideo_b200.default_config(geometric_verification=2, max_batch=4)) as c:
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        c.match_frames_bgr8(frames)
        ver = c.get_verification(0, len(frames))
        dec = c.get_decisions(0, len(frames))
        assert c.pool_pages_device_view()[2:] == (0, 0)
    seen = set()
    for f in range(len(frames)):
        ki, kf, dd = oracle.orb_detect_and_compute(oracle.gray_from_bgr(frames[f]))
        idx, dist = oracle.bf_knn_hamming(dd, pool, 30)
        want = ph.decide_frame(idx, dist.astype(np.float32), offs, kf[:, :2], pool_pts, frames[f], pages)
        assert ver[f]["survivors"] == [tuple(int(v) for v in x) for x in want["survivors"]]
        assert dec[f]["image"] == want["image"], f"frame {f}"
        assert [p for p, _ in dec[f]["rated"]] == [p for p, _ in want["rated"]]
        for (_, s), (_, ws) in zip(dec[f]["rated"], want["rated"]):
            assert abs(float(s) - float(ws)) <= 1e-6
        seen.add(dec[f]["image"])
    assert len({pages[i].shape for i in seen if i >= 0}) == 3          # pages of all three sizes were decided for
