"""Pins the prefilter restatement (oracle/area_oracle.c: INTER_AREA resize + similarity, image_utils.rs:8-27) against cv2."""
import json
import os
import zlib

import numpy as np
import pytest

import oracle
import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def test_small_size():
    g = json.load(open(os.path.join(GOLD, "area.json")))
    assert list(oracle.small_size(1920, 1080)) == g["small_size_1920x1080"] == [461, 259]
    assert list(oracle.small_size(2001, 1125)) == g["small_size_2001x1125"]


@pytest.mark.skipif(cv2 is None, reason="synthetic generator needs cv2")
def test_small_images_equal_golden():
    g = json.load(open(os.path.join(GOLD, "area.json")))
    smalls = {}
    for f in (3, 4, 5):
        smalls[f] = oracle.to_small_image(synth.make_frame(f, 50))
        assert crc(smalls[f]) == g["crc"][str(f)]
    page = np.repeat(synth.make_page(2)[:, :, None], 3, axis=2)
    assert crc(oracle.to_small_image(page)) == g["crc"]["page2"]
    for key, want in g["similarity"].items():
        a, b = (int(x) for x in key.split("-"))
        assert oracle.similarity(smalls[a], smalls[b]) == np.float32(want)
    assert oracle.similarity(smalls[3], smalls[3]) == np.float32(1.0)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_area_resize_equals_cv2_live():
    rng = np.random.default_rng(3)
    for (h, w) in ((1080, 1920), (720, 1280), (333, 517), (1125, 2001)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        dw, dh = oracle.small_size(w, h)
        assert np.array_equal(oracle.to_small_image(img), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA))
    a = rng.integers(0, 256, (259, 461, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (259, 461, 3), dtype=np.uint8)
    err = cv2.norm(a, b, cv2.NORM_L2)
    want = np.float32(1.0) - np.float32(err) / np.sqrt(np.float32(255.0 * 255.0 * 3.0) * np.float32(259 * 461), dtype=np.float32)
    assert oracle.similarity(a, b) == want
