"""CPU oracle for the slideo hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package, and only as the checker.  The product (``slideo_b200``) never imports it.

Two tiers:
  * ``liboracle.so``  -- plain-C restatement of OpenCV's ORB / brute-force k-NN / the reference's vote
                         (orb_oracle.c, bf_oracle.c; each function cites the reference file:line it follows)
  * ``cv2_oracle``    -- the same path through the real OpenCV (cv2 4.13.0), which is the third-party
                         library the reference's arithmetic lives in (reference pins 4.5.2).  It pins the
                         restatement; it is also the CPU baseline that bench.py times.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# reference ORB parameters: crates/matching-opencv/src/feature_extractor.rs:13-23
ORB_REF = dict(nfeatures=2000, scale_factor=1.2, nlevels=8, edge_threshold=62, patch_size=62, fast_threshold=20)
KNN_K = 30            # crates/matching-opencv/src/lib.rs:266
VOTE_RATIO = 1.05     # crates/matching-opencv/src/lib.rs:275


def build() -> str:
    """Compile liboracle.so (gcc) if missing or stale."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "bf_oracle.c", "ransac_oracle.c", "area_oracle.c", "sift_oracle.c", "libm_port.c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        c_i32p = ctypes.POINTER(ctypes.c_int32)
        c_f32p = ctypes.POINTER(ctypes.c_float)
        ci, cf = ctypes.c_int, ctypes.c_float
        L.orb_gray_from_bgr.argtypes = [c_u8p, ci, ci, ci, c_u8p]
        L.orb_level_scale.argtypes = [ci, cf]
        L.orb_level_scale.restype = cf
        L.orb_level_size.argtypes = [ci, ci, ci, cf, c_i32p, c_i32p]
        L.orb_level_quota.argtypes = [ci, ci, cf, c_i32p]
        L.orb_resize_linear_exact.argtypes = [c_u8p, ci, ci, c_u8p, ci, ci]
        L.orb_fast_score_map.argtypes = [c_u8p, ci, ci, ci, c_u8p]
        L.orb_fast_nms.argtypes = [c_u8p, ci, ci, c_i32p, c_i32p, c_i32p, ci]
        L.orb_fast_nms.restype = ci
        L.orb_fast_atan2.argtypes = [cf, cf]
        L.orb_fast_atan2.restype = cf
        L.orb_blur7.argtypes = [c_u8p, ci, ci, c_u8p]
        L.orb_pattern.argtypes = [ci, ci, c_i32p]
        L.orb_detect_and_compute.argtypes = [c_u8p, ci, ci, ci, cf, ci, ci, ci, ci, c_i32p, c_f32p, c_u8p, ci]
        L.orb_detect_and_compute.restype = ci
        L.bf_knn_hamming.argtypes = [c_u8p, ci, c_u8p, ci, ci, c_i32p, c_i32p]
        L.bf_knn_l2.argtypes = [c_f32p, ci, c_f32p, ci, ci, ci, c_i32p, c_f32p]
        L.bf_vote.argtypes = [c_i32p, c_f32p, ci, ci, c_i32p, ci, c_i32p]
        L.bf_vote.restype = ci
        c_dp = ctypes.POINTER(ctypes.c_double)
        L.ransac_affine_partial.argtypes = [c_f32p, c_f32p, ci, ctypes.c_double, ci, ctypes.c_double, ci, c_u8p, c_dp, c_i32p]
        L.ransac_affine_partial.restype = ci
        L.ransac_sample_sequence.argtypes = [ci, ci, c_i32p]
        L.ransac_update_num_iters.argtypes = [ctypes.c_double, ctypes.c_double, ci, ci]
        L.ransac_update_num_iters.restype = ci
        L.area_small_size.argtypes = [ci, ci, c_i32p, c_i32p]
        L.area_resize_u8.argtypes = [c_u8p, ci, ci, ci, ci, c_u8p, ci, ci, ci]
        L.area_similarity.argtypes = [c_u8p, c_u8p, ci, ci, ci]
        L.area_similarity.restype = cf
        c_dp2 = ctypes.POINTER(ctypes.c_double)
        L.sift_gauss_ksize.argtypes = [ctypes.c_double]
        L.sift_gauss_ksize.restype = ci
        L.sift_gauss_taps.argtypes = [ctypes.c_double, ci, c_f32p]
        L.sift_gauss_blur.argtypes = [c_f32p, ci, ci, c_f32p, ci, c_f32p]
        L.sift_upsample2.argtypes = [c_u8p, ci, ci, c_f32p]
        L.sift_exp32f.argtypes = [cf]
        L.sift_exp32f.restype = cf
        L.sift_fast_atan2.argtypes = [cf, cf]
        L.sift_fast_atan2.restype = cf
        L.sift_magnitude.argtypes = [cf, cf]
        L.sift_magnitude.restype = cf
        L.sift_num_octaves.argtypes = [ci, ci]
        L.sift_num_octaves.restype = ci
        L.sift_layer_sigmas.argtypes = [c_dp2]
        L.sift_pyramid_image.argtypes = [c_u8p, ci, ci, ci, ci, ci, c_f32p, c_i32p, c_i32p]
        L.sift_pyramid_image.restype = ci
        L.sift_detect_and_compute.argtypes = [c_u8p, ci, ci, c_f32p, c_i32p, c_f32p, ci]
        L.sift_detect_and_compute.restype = ci
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


# ----------------------------------------------------------------------------- ORB stages
def gray_from_bgr(bgr: np.ndarray) -> np.ndarray:
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().orb_gray_from_bgr(_p(bgr, ctypes.c_uint8), w, h, w * 3, _p(out, ctypes.c_uint8))
    return out


def level_sizes(w: int, h: int, nlevels: int = 8, scale_factor: float = 1.2):
    out = []
    for l in range(nlevels):
        lw, lh = ctypes.c_int32(), ctypes.c_int32()
        lib().orb_level_size(w, h, l, scale_factor, ctypes.byref(lw), ctypes.byref(lh))
        out.append((lw.value, lh.value))
    return out


def level_quota(nfeatures: int = 2000, nlevels: int = 8, scale_factor: float = 1.2):
    q = np.zeros(nlevels, np.int32)
    lib().orb_level_quota(nfeatures, nlevels, scale_factor, _p(q, ctypes.c_int32))
    return q.tolist()


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = _u8(src)
    out = np.empty((dh, dw), np.uint8)
    lib().orb_resize_linear_exact(_p(src, ctypes.c_uint8), src.shape[1], src.shape[0], _p(out, ctypes.c_uint8), dw, dh)
    return out


def fast_score_map(img: np.ndarray, thr: int = 20) -> np.ndarray:
    img = _u8(img)
    out = np.empty_like(img)
    lib().orb_fast_score_map(_p(img, ctypes.c_uint8), img.shape[1], img.shape[0], thr, _p(out, ctypes.c_uint8))
    return out


def fast_keypoints(img: np.ndarray, thr: int = 20):
    """(x, y, score) after 3x3 NMS, raster order."""
    sm = fast_score_map(img, thr)
    cap = img.size // 4 + 16
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orb_fast_nms(_p(sm, ctypes.c_uint8), img.shape[1], img.shape[0], _p(xs, ctypes.c_int32),
                           _p(ys, ctypes.c_int32), _p(sc, ctypes.c_int32), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def blur7(img: np.ndarray) -> np.ndarray:
    img = _u8(img)
    out = np.empty_like(img)
    lib().orb_blur7(_p(img, ctypes.c_uint8), img.shape[1], img.shape[0], _p(out, ctypes.c_uint8))
    return out


def pattern(half: int = 31, npoints: int = 512) -> np.ndarray:
    out = np.empty((npoints, 2), np.int32)
    lib().orb_pattern(half, npoints, _p(out, ctypes.c_int32))
    return out


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orb_fast_atan2(y, x))


def orb_detect_and_compute(gray: np.ndarray, nfeatures=2000, scale_factor=1.2, nlevels=8, edge_threshold=62,
                           patch_size=62, fast_threshold=20, cap=65536):
    """Restated ORB on an 8-bit gray image.  Canonical order (octave, y, x).

    Returns (kp_i [n,4] int32 {x_level,y_level,octave,score}, kp_f [n,4] f32 {pt.x,pt.y,size,angle}, desc [n,32] u8).
    """
    gray = _u8(gray)
    h, w = gray.shape
    kp_i = np.empty((cap, 4), np.int32)
    kp_f = np.empty((cap, 4), np.float32)
    desc = np.empty((cap, 32), np.uint8)
    n = lib().orb_detect_and_compute(_p(gray, ctypes.c_uint8), w, h, nfeatures, scale_factor, nlevels,
                                     edge_threshold, patch_size, fast_threshold, _p(kp_i, ctypes.c_int32),
                                     _p(kp_f, ctypes.c_float), _p(desc, ctypes.c_uint8), cap)
    if n < 0:
        raise RuntimeError(f"orb oracle failed ({n})")
    return kp_i[:n].copy(), kp_f[:n].copy(), desc[:n].copy()


# ----------------------------------------------------------------------------- matcher + vote
def bf_knn_hamming(q: np.ndarray, t: np.ndarray, k: int = KNN_K):
    q, t = _u8(q), _u8(t)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.int32)
    lib().bf_knn_hamming(_p(q, ctypes.c_uint8), len(q), _p(t, ctypes.c_uint8), len(t), k, _p(idx, ctypes.c_int32),
                         _p(dist, ctypes.c_int32))
    return idx, dist


def bf_knn_l2(q: np.ndarray, t: np.ndarray, k: int = KNN_K):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.float32)
    lib().bf_knn_l2(_p(q, ctypes.c_float), len(q), _p(t, ctypes.c_float), len(t), q.shape[1], k,
                    _p(idx, ctypes.c_int32), _p(dist, ctypes.c_float))
    return idx, dist


def vote(idx: np.ndarray, dist: np.ndarray, page_offsets):
    """Reference vote (lib.rs:268-282).  Returns (best_page or -1, votes_of_best, votes[P])."""
    idx = np.ascontiguousarray(idx, np.int32)
    dist = np.ascontiguousarray(dist, np.float32)
    po = np.ascontiguousarray(page_offsets, np.int32)
    votes = np.zeros(len(po) - 1, np.int32)
    best = lib().bf_vote(_p(idx, ctypes.c_int32), _p(dist, ctypes.c_float), idx.shape[0], idx.shape[1],
                         _p(po, ctypes.c_int32), len(po) - 1, _p(votes, ctypes.c_int32))
    return best, (int(votes[best]) if best >= 0 else 0), votes


def match_frame(frame_desc: np.ndarray, page_descs, k: int = KNN_K):
    """Whole matcher stage for one frame: pooled BF k-NN + vote -> (best_page, votes, votes[P])."""
    offs = np.zeros(len(page_descs) + 1, np.int32)
    offs[1:] = np.cumsum([len(d) for d in page_descs])
    pool = np.concatenate([_u8(d).reshape(-1, 32) for d in page_descs], 0) if offs[-1] else np.zeros((0, 32), np.uint8)
    idx, dist = bf_knn_hamming(frame_desc, pool, k)
    return vote(idx, dist.astype(np.float32), offs)


# ----------------------------------------------------------------------------- geometric verification (lib.rs:284-333)
RANSAC_THRESHOLD, RANSAC_MAX_ITERS, RANSAC_CONFIDENCE = 3.0, 2000, 0.99   # image_utils.rs:52
TOP_SLIDES, TOP_RATED, MIN_RATING, MIN_RATING_FRACTION = 40, 10, 50.0, 0.2  # lib.rs:295, 330, 333


def ransac_affine_partial(pts_from: np.ndarray, pts_to: np.ndarray):
    """cv::estimateAffinePartial2D(from, to, inliers, RANSAC, 3.0, 2000, 0.99, 10) -> (rating, mask, ransac_model[6], iters)."""
    a = np.ascontiguousarray(pts_from, np.float32).reshape(-1, 2)
    b = np.ascontiguousarray(pts_to, np.float32).reshape(-1, 2)
    n = len(a)
    mask = np.zeros(max(n, 1), np.uint8)
    model = np.zeros(6, np.float64)
    it = ctypes.c_int32()
    good = lib().ransac_affine_partial(_p(a, ctypes.c_float), _p(b, ctypes.c_float), n, RANSAC_THRESHOLD, RANSAC_MAX_ITERS,
                                       RANSAC_CONFIDENCE, 0, _p(mask, ctypes.c_uint8), _p(model, ctypes.c_double), ctypes.byref(it))
    return good, mask[:n].copy(), model, it.value


def verify_frame(idx: np.ndarray, dist: np.ndarray, page_offsets, frame_pts: np.ndarray, pool_pts: np.ndarray):
    """The reference's ranking + geometric gate for one frame (lib.rs:268-333) on exact k-NN rows.

    idx/dist: [nq, k] rows (oracle order); frame_pts [nq, 2], pool_pts [Nt, 2] keypoint coordinates (KeyPoint.pt).
    Returns dict(cand=[(page, votes, rating)] in ranking order (<= 40), survivors=[(page, rating)] (<= 10)).
    Ties (votes, rating) resolve to the lower page index (the reference's HashMap order is arbitrary)."""
    po = np.ascontiguousarray(page_offsets, np.int64)
    npages = len(po) - 1
    by_page = [[] for _ in range(npages)]
    dist = np.asarray(dist, np.float32)
    for q in range(idx.shape[0]):
        if idx[q, 0] < 0:
            continue
        lim = np.float32(dist[q, 0]) * np.float32(VOTE_RATIO)
        for j in range(idx.shape[1]):
            g = idx[q, j]
            if g >= 0 and dist[q, j] < lim:
                by_page[int(np.searchsorted(po, g, side="right") - 1)].append((q, int(g)))
    order = sorted((p for p in range(npages) if by_page[p]), key=lambda p: (-len(by_page[p]), p))[:TOP_SLIDES]
    cand = []
    for p in order:
        m = by_page[p]
        fr = pool_pts[[g for _, g in m]]          # slide keypoints  (lib.rs:299)
        to = frame_pts[[q for q, _ in m]]         # frame keypoints  (lib.rs:300)
        rating, _, _, _ = ransac_affine_partial(fr, to)
        cand.append((p, len(m), rating))
    ranked = sorted(cand, key=lambda c: -c[2])[:TOP_RATED]       # stable: ties keep the vote order (Rust sort_by is stable)
    best = float(ranked[0][2]) if ranked else 0.0
    surv = [(p, r) for p, _, r in ranked if r > MIN_RATING and (float(r) / best if best else 0.0) > MIN_RATING_FRACTION]
    return dict(cand=cand, survivors=surv)


# ----------------------------------------------------------------------------- changed-frame prefilter (video_capture.rs:60-103)
CHANGED_THRESHOLD = np.float32(0.98)   # video_capture.rs:98


def small_size(w: int, h: int):
    sw, sh = ctypes.c_int32(), ctypes.c_int32()
    lib().area_small_size(w, h, ctypes.byref(sw), ctypes.byref(sh))
    return sw.value, sh.value


def to_small_image(img: np.ndarray) -> np.ndarray:
    """image_utils.rs:8-19: cv::resize(img, small_size, INTER_AREA) for 8-bit images (1 or 3 interleaved channels)."""
    img = _u8(img)
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    dw, dh = small_size(w, h)
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().area_resize_u8(_p(img, ctypes.c_uint8), w, h, img.strides[0], cn, _p(out, ctypes.c_uint8), dw, dh, 0)
    return out


def similarity(a: np.ndarray, b: np.ndarray) -> np.float32:
    """image_utils.rs:21-27."""
    a, b = _u8(a), _u8(b)
    cn = 1 if a.ndim == 2 else a.shape[2]
    return np.float32(lib().area_similarity(_p(a, ctypes.c_uint8), _p(b, ctypes.c_uint8), a.shape[1], a.shape[0], cn))


def mark_similar(frames):
    """MarkSimilarIter: (changed[n], similarity[n]) for consecutive sampled frames."""
    last, ch, sims = None, [], []
    for f in frames:
        small = to_small_image(f)
        s = similarity(last, small) if last is not None else np.float32(0.0)
        last = small
        sims.append(s)
        ch.append(bool(s < CHANGED_THRESHOLD))
    return np.array(ch, bool), np.array(sims, np.float32)


# ----------------------------------------------------------------------------- SIFT (north_star variant; cv2.SIFT_create() defaults)
def sift_layer_sigmas():
    """[sig_diff of the initial image, sig[1..5] of buildGaussianPyramid] (nOctaveLayers 3, sigma 1.6)."""
    s = np.zeros(6, np.float64)
    lib().sift_layer_sigmas(_p(s, ctypes.c_double))
    return s


def sift_gauss_taps(sigma: float) -> np.ndarray:
    n = lib().sift_gauss_ksize(sigma)
    t = np.zeros(n, np.float32)
    lib().sift_gauss_taps(sigma, n, _p(t, ctypes.c_float))
    return t


def sift_gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(img_f32, (0, 0), sigma) restated."""
    img = np.ascontiguousarray(img, np.float32)
    t = sift_gauss_taps(sigma)
    out = np.empty_like(img)
    lib().sift_gauss_blur(_p(img, ctypes.c_float), img.shape[1], img.shape[0], _p(t, ctypes.c_float), len(t), _p(out, ctypes.c_float))
    return out


def sift_upsample2(gray: np.ndarray) -> np.ndarray:
    gray = _u8(gray)
    out = np.empty((2 * gray.shape[0], 2 * gray.shape[1]), np.float32)
    lib().sift_upsample2(_p(gray, ctypes.c_uint8), gray.shape[1], gray.shape[0], _p(out, ctypes.c_float))
    return out


def sift_exp(x: np.ndarray) -> np.ndarray:
    f = lib().sift_exp32f
    return np.array([f(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32)


def sift_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    f = lib().sift_fast_atan2
    return np.array([f(float(a), float(b)) for a, b in zip(np.asarray(y, np.float32).ravel(), np.asarray(x, np.float32).ravel())], np.float32)


def sift_magnitude(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    f = lib().sift_magnitude
    return np.array([f(float(a), float(b)) for a, b in zip(np.asarray(x, np.float32).ravel(), np.asarray(y, np.float32).ravel())], np.float32)


def sift_pyramid_image(gray: np.ndarray, kind: int, octave: int, layer: int) -> np.ndarray:
    """kind 0: Gaussian pyramid image, 1: DoG image (octave 0 = the doubled image)."""
    gray = _u8(gray)
    w, h = ctypes.c_int32(), ctypes.c_int32()
    if lib().sift_pyramid_image(_p(gray, ctypes.c_uint8), gray.shape[1], gray.shape[0], kind, octave, layer, None, ctypes.byref(w), ctypes.byref(h)):
        raise ValueError("octave out of range")
    out = np.empty((h.value, w.value), np.float32)
    lib().sift_pyramid_image(_p(gray, ctypes.c_uint8), gray.shape[1], gray.shape[0], kind, octave, layer, _p(out, ctypes.c_float), ctypes.byref(w), ctypes.byref(h))
    return out


def sift_detect_and_compute(gray: np.ndarray, cap: int = 65536):
    """Restated cv2.SIFT_create().detectAndCompute on an 8-bit gray image, in OpenCV's output order.

    Returns (kp_f [n,5] f32 {pt.x, pt.y, size, angle, response}, octave [n] int32 (packed), desc [n,128] f32)."""
    gray = _u8(gray)
    h, w = gray.shape
    kp_f = np.empty((cap, 5), np.float32)
    octv = np.empty(cap, np.int32)
    desc = np.empty((cap, 128), np.float32)
    n = lib().sift_detect_and_compute(_p(gray, ctypes.c_uint8), w, h, _p(kp_f, ctypes.c_float), _p(octv, ctypes.c_int32), _p(desc, ctypes.c_float), cap)
    if n > cap:
        return sift_detect_and_compute(gray, n)
    return kp_f[:n].copy(), octv[:n].copy(), desc[:n].copy()
