"""Regenerates the golden vectors under tests/golden/ from the REAL OpenCV (cv2) -- the third-party library the
reference's hot-path arithmetic lives in (crates/matching-opencv/Cargo.toml:8; reference pins OpenCV 4.5.2, this
image has cv2 4.13.0) -- called with the reference's literals (feature_extractor.rs:13-23, lib.rs:266,275).

Run in the build container:  python tests/golden/make_golden.py
  * reference_fixtures.json : per reference fixture PNG (data/matchings/test1/*.png, read from /root/reference,
                              never copied): keypoint count + CRC32 of cv2's canonical keypoints / descriptors, and
                              the frame->slide vote anchors of SURVEY.md section 4.
  * synth_orb.npz           : cv2 ORB outputs on seeded synthetic images (the images themselves are regenerated from
                              synth.py by the tests, only cv2's outputs are stored).
  * bf_knn.npz              : cv2.BFMatcher knnMatch outputs (Hamming + L2) on small seeded pools with planted ties,
                              and the reference vote computed on cv2's DMatch rows.
"""
import json
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from oracle import cv2_oracle as co  # noqa: E402

REF_FIX = "/root/reference/data/matchings/test1"


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def synth_images():
    rng = np.random.default_rng(42)
    small = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    return {
        "page1": synth.make_page(1),
        "frame7": synth.make_frame(7, 50),
        "texture": cv2.resize(small, (640, 480), interpolation=cv2.INTER_CUBIC),
    }


def main():
    out = {"opencv_version": cv2.__version__, "fixtures": {}, "votes": {}}
    descs = {}
    for name in ("1-frame", "2-frame", "3-frame", "1-slide", "3-slide"):
        path = os.path.join(REF_FIX, name + ".png")
        if name.endswith("slide"):
            img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)      # lib.rs:98
            img3 = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)      # lib.rs:104 (gray replicated)
        else:
            img3 = cv2.imread(path, cv2.IMREAD_COLOR)          # frames arrive BGR (video_capture.rs:45-53)
        ki, kf, d = co.orb_canonical(img3)
        descs[name] = d
        out["fixtures"][name] = {"shape": list(img3.shape), "n": len(ki), "crc_kp_i": crc(ki), "crc_kp_f": crc(kf), "crc_desc": crc(d)}
    pages = [descs["1-slide"], descs["3-slide"]]
    for f in ("1-frame", "2-frame", "3-frame"):
        best, votes, allv = co.match_frame_bf(descs[f], pages)
        out["votes"][f] = {"best": best, "votes": votes, "all": [int(v) for v in allv]}
    # ORB geometry constants
    out["level_sizes_1920x1080"] = [[1920, 1080], [1600, 900], [1333, 750], [1111, 625], [926, 521], [772, 434], [643, 362], [536, 301]]
    out["quota_2000"] = [434, 362, 302, 251, 209, 175, 145, 122]
    out["quota_500"] = [109, 90, 75, 63, 52, 44, 36, 31]
    out["pattern_crc32"] = 0xDD317DF4
    with open(os.path.join(HERE, "reference_fixtures.json"), "w") as fh:
        json.dump(out, fh, indent=1)

    arrs = {}
    for name, img in synth_images().items():
        for nf in (2000, 500):
            ki, kf, d = co.orb_canonical(img, nfeatures=nf)
            arrs[f"{name}_{nf}_kp_i"] = ki
            arrs[f"{name}_{nf}_kp_f"] = kf
            arrs[f"{name}_{nf}_desc"] = d
    g = cv2.cvtColor(synth_images()["frame7"], cv2.COLOR_BGR2GRAY)
    arrs["frame7_gray_crc"] = np.array([crc(g)], np.int64)
    lvl1 = cv2.resize(synth_images()["page1"], (1668, 938), interpolation=cv2.INTER_LINEAR_EXACT)
    arrs["page1_level1_crc"] = np.array([crc(lvl1)], np.int64)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    arrs["page1_blur_crc"] = np.array([crc(cv2.sepFilter2D(synth_images()["page1"], cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))], np.int64)
    fast = cv2.FastFeatureDetector_create(20, True).detect(synth_images()["texture"], None)
    arrs["texture_fast"] = np.array(sorted((int(p.pt[1]), int(p.pt[0]), int(p.response)) for p in fast), np.int32)
    np.savez_compressed(os.path.join(HERE, "synth_orb.npz"), **arrs)

    rng = np.random.default_rng(123)
    pages = [synth.hamming_pool(n, seed=300 + i, dup_frac=0.05) for i, n in enumerate((250, 0, 180, 170))]
    pages[3][:60] = pages[0][:60]                    # cross-page duplicates -> ties at the k boundary
    pool = np.concatenate(pages)
    q = synth.hamming_queries(pool, 96, seed=301, near_frac=0.75)
    q[0] = pool[5]                                   # exact hit: best distance 0 -> casts no vote
    idx, dist = co.bf_knn_hamming(q, [p for p in pages if len(p)], 30)   # cv2 rejects empty Mats; offsets are unchanged
    m = cv2.BFMatcher(cv2.NORM_HAMMING)
    m.add([p for p in pages if len(p)])
    rows = m.knnMatch(q, 30)
    nonempty = [i for i, p in enumerate(pages) if len(p)]
    v = co.vote_rows(rows, len(nonempty))
    votes = np.zeros(len(pages), np.int64)
    votes[nonempty] = v
    t = np.rint(rng.normal(20, 30, (400, 128)).clip(0, 255)).astype(np.float32)
    t[300:320] = t[10:30]                            # exact duplicates
    ql = np.rint((t[rng.integers(0, 400, 24)] + rng.normal(0, 6, (24, 128))).clip(0, 255)).astype(np.float32)
    lidx, ldist = co.bf_knn_l2(ql, [t[:150], t[150:]], 30)
    np.savez_compressed(os.path.join(HERE, "bf_knn.npz"), ham_q=q, ham_pages_len=np.array([len(p) for p in pages]), ham_pool=pool,
                        ham_idx=idx, ham_dist=dist, ham_votes=votes, l2_q=ql, l2_t=t, l2_idx=lidx, l2_dist=ldist)
    print("golden vectors written;", {k: v for k, v in out["votes"].items()})


if __name__ == "__main__":
    main()


def make_ransac_golden():
    """tests/golden/ransac.npz: cv2.estimateAffinePartial2D (RANSAC, 3.0, 2000, 0.99, 10) inlier masks -- the call of
    crates/matching-opencv/src/image_utils.rs:52 -- on seeded correspondence sets."""
    rng = np.random.default_rng(77)
    sets = {}
    for i, (n, frac, sig) in enumerate([(3, 1.0, 0.5), (7, 0.6, 1.0), (60, 0.3, 2.0), (200, 0.05, 2.0), (200, 0.0, 1.0),
                                        (1500, 0.4, 1.5), (2500, 0.9, 2.5), (40, 0.5, 3.0)]):
        fr = rng.uniform(0, 2000, (n, 2)).astype(np.float32)
        s = rng.uniform(0.9, 1.1)
        a = rng.uniform(-0.05, 0.05)
        R = np.array([[s * np.cos(a), -s * np.sin(a)], [s * np.sin(a), s * np.cos(a)]])
        to = (fr @ R.T + rng.uniform(-20, 20, 2)).astype(np.float32)
        out = rng.random(n) > frac
        to[out] = rng.uniform(0, 2000, (int(out.sum()), 2)).astype(np.float32)
        to += rng.normal(0, sig, to.shape).astype(np.float32)
        M, inl = cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ransacReprojThreshold=3.0, maxIters=2000, confidence=0.99,
                                             refineIters=10)
        sets[f"from{i}"] = fr
        sets[f"to{i}"] = to
        sets[f"mask{i}"] = inl.ravel().astype(np.uint8) if inl is not None else np.zeros(n, np.uint8)
    sets["n_sets"] = np.array([8])
    np.savez_compressed(os.path.join(HERE, "ransac.npz"), **sets)


if __name__ == "__main__":
    make_ransac_golden()


def make_area_golden():
    """tests/golden/area.json: cv2.resize(INTER_AREA) to the reference's small size (image_utils.rs:8-19) as CRC32, and
    compute_similarity (image_utils.rs:21-27) values from cv2.norm, on seeded synthetic frames."""
    out = {"small_size_1920x1080": [461, 259], "crc": {}, "similarity": {}}
    smalls = {}
    for f in (3, 4, 5):
        fr = synth.make_frame(f, 50)
        sm = cv2.resize(fr, (461, 259), interpolation=cv2.INTER_AREA)
        smalls[f] = sm
        out["crc"][str(f)] = crc(sm)
    pg = cv2.cvtColor(synth.make_page(2), cv2.COLOR_GRAY2BGR)
    fac = np.sqrt(np.float32(120000) / np.float32(2001 * 1125), dtype=np.float32)
    dw, dh = int(np.float32(2001) * fac), int(np.float32(1125) * fac)
    out["small_size_2001x1125"] = [dw, dh]
    out["crc"]["page2"] = crc(cv2.resize(pg, (dw, dh), interpolation=cv2.INTER_AREA))
    for a, b in ((3, 4), (4, 5), (3, 3)):
        err = cv2.norm(smalls[a], smalls[b], cv2.NORM_L2)
        sim = np.float32(1.0) - np.float32(err) / np.sqrt(np.float32(255.0 * 255.0 * 3.0) * np.float32(259 * 461), dtype=np.float32)
        out["similarity"][f"{a}-{b}"] = float(sim)
    with open(os.path.join(HERE, "area.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    make_area_golden()
