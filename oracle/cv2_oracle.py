"""The hot path through the real OpenCV (cv2) -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

OpenCV is the third-party library all of the reference's hot-path arithmetic lives in
(crates/matching-opencv/Cargo.toml:8 -> opencv 0.52 -> system OpenCV 4.5.2; here: cv2 4.13.0).
Parameters are the reference's literals:
  ORB      : crates/matching-opencv/src/feature_extractor.rs:13-23
  k = 30   : crates/matching-opencv/src/lib.rs:266
  vote     : crates/matching-opencv/src/lib.rs:268-282
  LSH index: crates/matching-opencv/src/flann.rs:15-23 (timing only: approximate + non-deterministic)
"""
from __future__ import annotations

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def available() -> bool:
    return cv2 is not None


def make_orb(nfeatures: int = 2000):
    return cv2.ORB_create(nfeatures, 1.2, 8, 62, 0, 2, cv2.ORB_FAST_SCORE, 62, 20)


def orb_canonical(img: np.ndarray, nfeatures: int = 2000, orb=None):
    """cv2 ORB, re-ordered canonically by (octave, y_level, x_level).

    Returns kp_i [n,4] int32 {x_level,y_level,octave,score}, kp_f [n,4] f32 {pt.x,pt.y,size,angle}, desc [n,32].
    """
    orb = orb or make_orb(nfeatures)
    kps, desc = orb.detectAndCompute(img, None)
    if desc is None:
        return np.zeros((0, 4), np.int32), np.zeros((0, 4), np.float32), np.zeros((0, 32), np.uint8)
    n = len(kps)
    kp_i = np.empty((n, 4), np.int32)
    kp_f = np.empty((n, 4), np.float32)
    for i, k in enumerate(kps):
        sc = np.float32(np.float64(np.float32(1.2)) ** k.octave)
        inv = np.float32(1.0) / sc
        kp_i[i] = (int(np.rint(np.float32(k.pt[0]) * inv)), int(np.rint(np.float32(k.pt[1]) * inv)), k.octave,
                   int(k.response))
        kp_f[i] = (k.pt[0], k.pt[1], k.size, k.angle)
    order = np.lexsort((kp_i[:, 0], kp_i[:, 1], kp_i[:, 2]))
    return kp_i[order], kp_f[order], desc[order]


def bf_knn_hamming(q: np.ndarray, page_descs, k: int = 30):
    """cv2.BFMatcher(NORM_HAMMING) with one train Mat per page.  Returns global idx [nq,k], dist [nq,k] (int)."""
    m = cv2.BFMatcher(cv2.NORM_HAMMING)
    m.add([np.ascontiguousarray(d) for d in page_descs])
    offs = np.zeros(len(page_descs) + 1, np.int64)
    offs[1:] = np.cumsum([len(d) for d in page_descs])
    rows = m.knnMatch(np.ascontiguousarray(q), k)
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.full((len(q), k), -1, np.int32)
    for i, r in enumerate(rows):
        for j, dm in enumerate(r):
            idx[i, j] = offs[dm.imgIdx] + dm.trainIdx
            dist[i, j] = int(dm.distance)
    return idx, dist


def bf_knn_l2(q: np.ndarray, page_descs, k: int = 30):
    m = cv2.BFMatcher(cv2.NORM_L2)
    m.add([np.ascontiguousarray(d, np.float32) for d in page_descs])
    offs = np.zeros(len(page_descs) + 1, np.int64)
    offs[1:] = np.cumsum([len(d) for d in page_descs])
    rows = m.knnMatch(np.ascontiguousarray(q, np.float32), k)
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.full((len(q), k), -1, np.float32)
    for i, r in enumerate(rows):
        for j, dm in enumerate(r):
            idx[i, j] = offs[dm.imgIdx] + dm.trainIdx
            dist[i, j] = dm.distance
    return idx, dist


def vote_rows(rows, npages: int):
    """The reference's vote loop verbatim in behaviour (lib.rs:268-282) on cv2 DMatch rows."""
    votes = np.zeros(npages, np.int64)
    for r in rows:
        if not r:
            continue
        best = np.float32(r[0].distance)
        lim = best * np.float32(1.05)
        for dm in r:
            if np.float32(dm.distance) < lim:
                votes[dm.imgIdx] += 1
    return votes


def match_frame_bf(frame_desc: np.ndarray, page_descs, k: int = 30):
    """(best_page or -1, votes_of_best, votes[P]) with the exact matcher; ties -> lowest page index."""
    m = cv2.BFMatcher(cv2.NORM_HAMMING)
    m.add([np.ascontiguousarray(d) for d in page_descs])
    votes = vote_rows(m.knnMatch(np.ascontiguousarray(frame_desc), k), len(page_descs))
    best = int(np.argmax(votes)) if votes.max(initial=0) > 0 else -1
    return best, (int(votes[best]) if best >= 0 else 0), votes


def make_lsh_matcher(page_descs):
    """The matcher the reference really builds (flann.rs:15-23, 64-71).  Approximate; timing only."""
    m = cv2.FlannBasedMatcher(dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1),
                              dict(checks=32, eps=0.0, sorted=True))
    m.add([np.ascontiguousarray(d) for d in page_descs])
    m.train()
    return m
