"""Pins the ORB restatement (oracle/orb_oracle.c) against the real OpenCV:
  * committed golden vectors produced by cv2 (tests/golden/make_golden.py),
  * the reference's own fixture PNGs (data/matchings/test1) when /root/reference is mounted,
  * cv2 live when importable (stage by stage: gray, resize, FAST, blur; end to end).
The reference itself holds no test or golden vector for this path -- these are the pins."""
import json
import os
import zlib

import numpy as np
import pytest

import oracle
import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixtures")   # copies of the reference's data/matchings/test1/*.png (test infrastructure)

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "synth_orb.npz")), json.load(open(os.path.join(GOLD, "reference_fixtures.json")))


def _synth_images():
    rng = np.random.default_rng(42)
    small = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    if cv2 is None:
        pytest.skip("synthetic generator needs cv2")
    return {"page1": synth.make_page(1), "frame7": synth.make_frame(7, 50),
            "texture": cv2.resize(small, (640, 480), interpolation=cv2.INTER_CUBIC)}


def test_geometry_constants(gold):
    _, j = gold
    assert [list(s) for s in oracle.level_sizes(1920, 1080)] == j["level_sizes_1920x1080"]
    assert oracle.level_sizes(2001, 1125)[1] == (1668, 938)          # fp32 reciprocal, not division
    assert oracle.level_quota(2000) == j["quota_2000"]
    assert oracle.level_quota(500) == j["quota_500"]
    pat = oracle.pattern()
    assert crc(pat.astype("<i4")) == j["pattern_crc32"]
    assert pat[:4].tolist() == [[21, 14], [-13, -26], [11, 16], [-17, 13]]


def test_fast_atan2_constants():
    # quadrant handling + the fp32 polynomial
    assert oracle.fast_atan2(0.0, 1.0) == 0.0
    assert abs(oracle.fast_atan2(1.0, 1.0) - 45.0) < 0.02
    assert abs(oracle.fast_atan2(1.0, -1.0) - 135.0) < 0.02
    assert abs(oracle.fast_atan2(-1.0, -1.0) - 225.0) < 0.02
    assert abs(oracle.fast_atan2(-1.0, 1.0) - 315.0) < 0.02
    assert abs(oracle.fast_atan2(1.0, 0.0) - 90.0) < 1e-4


@pytest.mark.parametrize("name", ["page1", "frame7", "texture"])
@pytest.mark.parametrize("nf", [2000, 500])
def test_orb_equals_golden(gold, name, nf):
    g, _ = gold
    img = _synth_images()[name]
    gray = oracle.gray_from_bgr(img) if img.ndim == 3 else img
    ki, kf, d = oracle.orb_detect_and_compute(gray, nfeatures=nf)
    assert np.array_equal(ki, g[f"{name}_{nf}_kp_i"])
    assert np.array_equal(kf.view(np.uint32), g[f"{name}_{nf}_kp_f"].view(np.uint32))
    assert np.array_equal(d, g[f"{name}_{nf}_desc"])


def test_stages_equal_golden(gold):
    g, _ = gold
    ims = _synth_images()
    assert crc(oracle.gray_from_bgr(ims["frame7"])) == int(g["frame7_gray_crc"][0])
    assert crc(oracle.resize_linear_exact(ims["page1"], 1668, 938)) == int(g["page1_level1_crc"][0])
    assert crc(oracle.blur7(ims["page1"])) == int(g["page1_blur_crc"][0])
    xs, ys, sc = oracle.fast_keypoints(ims["texture"])
    assert np.array_equal(np.stack([ys, xs, sc], 1).astype(np.int32), g["texture_fast"])


@pytest.mark.skipif(not os.path.isdir(REF_FIX) or cv2 is None, reason="reference fixtures not mounted")
@pytest.mark.parametrize("name", ["1-frame", "2-frame", "3-frame", "1-slide", "3-slide"])
def test_orb_on_reference_fixtures(gold, name):
    _, j = gold
    path = os.path.join(REF_FIX, name + ".png")
    if name.endswith("slide"):
        gray = cv2.imread(path, cv2.IMREAD_GRAYSCALE)                 # lib.rs:98
    else:
        gray = oracle.gray_from_bgr(cv2.imread(path, cv2.IMREAD_COLOR))
    ki, kf, d = oracle.orb_detect_and_compute(gray)
    f = j["fixtures"][name]
    assert (len(ki), crc(ki), crc(kf), crc(d)) == (f["n"], f["crc_kp_i"], f["crc_kp_f"], f["crc_desc"])


@pytest.mark.skipif(not os.path.isdir(REF_FIX) or cv2 is None, reason="reference fixtures not mounted")
def test_vote_anchors_on_reference_fixtures(gold):
    """frame1->slide1 2347 votes, frame3->slide3 2545 votes (BF variant of the reference pipeline)."""
    _, j = gold
    desc = {}
    for name in ("1-frame", "2-frame", "3-frame", "1-slide", "3-slide"):
        path = os.path.join(REF_FIX, name + ".png")
        gray = cv2.imread(path, 0) if name.endswith("slide") else oracle.gray_from_bgr(cv2.imread(path, 1))
        desc[name] = oracle.orb_detect_and_compute(gray)[2]
    for f in ("1-frame", "2-frame", "3-frame"):
        best, votes, allv = oracle.match_frame(desc[f], [desc["1-slide"], desc["3-slide"]])
        assert (best, votes, allv.tolist()) == (j["votes"][f]["best"], j["votes"][f]["votes"], j["votes"][f]["all"])
    assert j["votes"]["1-frame"]["votes"] == 2347 and j["votes"]["3-frame"]["votes"] == 2545


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_stages_equal_cv2_live():
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(oracle.gray_from_bgr(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    g = rng.integers(0, 256, (301, 403), dtype=np.uint8)
    for (dw, dh) in ((336, 251), (200, 150), (402, 300)):
        assert np.array_equal(oracle.resize_linear_exact(g, dw, dh), cv2.resize(g, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    assert np.array_equal(oracle.blur7(g), cv2.sepFilter2D(g, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))
    sm = cv2.resize(rng.integers(0, 256, (40, 50), dtype=np.uint8), (400, 320), interpolation=cv2.INTER_CUBIC)
    kps = cv2.FastFeatureDetector_create(20, True).detect(sm, None)
    want = sorted((int(p.pt[1]), int(p.pt[0]), int(p.response)) for p in kps)
    xs, ys, sc = oracle.fast_keypoints(sm)
    assert list(zip(ys.tolist(), xs.tolist(), sc.tolist())) == want


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_orb_equals_cv2_live_small():
    from oracle import cv2_oracle as co
    rng = np.random.default_rng(5)
    img = cv2.resize(rng.integers(0, 256, (48, 64), dtype=np.uint8), (512, 384), interpolation=cv2.INTER_CUBIC)
    ki, kf, d = co.orb_canonical(img)
    oi, of, od = oracle.orb_detect_and_compute(img)
    assert len(ki) > 100
    assert np.array_equal(ki, oi) and np.array_equal(kf.view(np.uint32), of.view(np.uint32)) and np.array_equal(d, od)
