"""K1-K7 parity: GPU ORB (through the C ABI) == the oracle restatement (orb_oracle.c, itself pinned bit-exact
against cv2 4.13 in tests/test_oracle_orb.py): keypoint set, scores, angles, coordinates and all 256 descriptor
bits, in canonical (octave, y, x) order.  Stage-level checks (pyramid, blur, FAST candidates) localise failures."""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


def _images():
    rng = np.random.default_rng(42)
    small = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    import cv2
    tex = cv2.resize(small, (640, 480), interpolation=cv2.INTER_CUBIC)
    return {
        "page": synth.make_page(1),                                   # 2001 x 1125 (odd width)
        "frame_bgr": synth.make_frame(7, 50),                          # 1920 x 1080 x 3
        "texture": tex,                                               # many corners
        "flat": np.full((400, 600), 128, np.uint8),                   # no keypoints at all
        "noise_odd": rng.integers(0, 256, (333, 517), dtype=np.uint8),  # ragged size, saturating candidate lists
    }


@pytest.mark.parametrize("name", ["page", "frame_bgr", "texture", "flat", "noise_odd"])
def test_orb_matches_oracle(ctx, name):
    img = _images()[name]
    gray = oracle.gray_from_bgr(img) if img.ndim == 3 else img
    oi, of, od = oracle.orb_detect_and_compute(gray)
    gi, gf, gd = ctx.extract_orb(img)
    # stage-level first: pyramid and blurred levels
    sizes = oracle.level_sizes(gray.shape[1], gray.shape[0])
    lvl = gray
    for l, (w, h) in enumerate(sizes):
        if l:
            lvl = oracle.resize_linear_exact(lvl, w, h)
        assert np.array_equal(ctx.debug_fetch(0, l), lvl), f"pyramid level {l}"
        assert np.array_equal(ctx.debug_fetch(1, l), oracle.blur7(lvl)), f"blurred level {l}"
    assert gi.shape == oi.shape, f"{len(gi)} vs {len(oi)} keypoints"
    assert np.array_equal(gi, oi)
    assert np.array_equal(gf.view(np.uint32), of.view(np.uint32)), "pt / size / angle differ bitwise"
    assert np.array_equal(gd, od)


def test_orb_nfeatures_500(ctx):
    """BASELINE config 1 uses ORB-500: quota [109,90,75,63,52,44,36,31]."""
    import slideo_b200
    img = synth.make_page(2)
    with slideo_b200.Context(slideo_b200.default_config(nfeatures=500)) as c:
        gi, gf, gd = c.extract_orb(img)
    oi, of, od = oracle.orb_detect_and_compute(img, nfeatures=500)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od) and np.array_equal(gf.view(np.uint32), of.view(np.uint32))


def test_fast_candidates_match_oracle(ctx):
    img = synth.make_page(4)
    ctx.extract_orb(img)
    sizes = oracle.level_sizes(img.shape[1], img.shape[0])
    lvl = img
    for l, (w, h) in enumerate(sizes):
        if l:
            lvl = oracle.resize_linear_exact(lvl, w, h)
        xs, ys, sc = oracle.fast_keypoints(lvl)
        keep = (xs >= 62) & (xs < w - 62) & (ys >= 62) & (ys < h - 62)
        want = np.sort((sc[keep].astype(np.uint32) << 24) | (ys[keep].astype(np.uint32) << 12) | xs[keep].astype(np.uint32))
        got = np.sort(ctx.debug_fetch(2, l))
        assert np.array_equal(got, want), f"level {l}"
