"""Golden vectors for the SIFT variant (north_star; SIFT is not in the reference -- SURVEY.md D5): outputs of the REAL
OpenCV, cv2.SIFT_create().detectAndCompute with its defaults, on seeded images that the tests regenerate themselves.

cv2 is run with IPP off and one thread: the default wheel routes hal::exp32f / magnitude32f through closed-source Intel
IPP (ippsExp_32f_A21, ippsMagnitude_32f) and its multi-threaded run is not even bit-deterministic in KeyPoint.angle
(measured: up to 57 of 2120 angles change between two runs); IPP off + one thread is OpenCV's own open-source code path
and is deterministic.

Run in the build container:  python tests/golden/make_golden_sift.py  ->  tests/golden/sift.npz
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import synth  # noqa: E402


def images():
    """Inputs are built with IPP off as well (cv2.resize / warpAffine results depend on it), state restored afterwards."""
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        return _images()
    finally:
        cv2.ipp.setUseIPP(ipp)


def _images():
    rng = np.random.default_rng(77)
    small = rng.integers(0, 256, (40, 56), dtype=np.uint8)
    return {
        "texture": cv2.resize(small, (333, 250), interpolation=cv2.INTER_CUBIC),       # odd width: blur remainder columns
        "page_crop": np.ascontiguousarray(synth.make_page(3)[100:500, 40:680]),        # text + shapes, 640 x 400
        "frame_crop": np.ascontiguousarray(cv2.cvtColor(synth.make_frame(5, 50), cv2.COLOR_BGR2GRAY)[300:620, 500:980]),
    }


def cv2_sift(gray):
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(False)
    kp, d = cv2.SIFT_create().detectAndCompute(gray, None)
    kf = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kp], np.float32).reshape(-1, 5)
    oc = np.array([k.octave for k in kp], np.int32)
    return kf, oc, (d if d is not None else np.zeros((0, 128), np.float32))


def main():
    out = {"opencv_version": np.array(cv2.__version__)}
    for name, img in images().items():
        kf, oc, d = cv2_sift(img)
        out[name + "_kp"] = kf
        out[name + "_octave"] = oc
        out[name + "_desc"] = d.astype(np.uint8)   # integer-valued 0..255
        print(name, img.shape, len(kf))
    np.savez_compressed(os.path.join(HERE, "sift.npz"), **out)


if __name__ == "__main__":
    main()
