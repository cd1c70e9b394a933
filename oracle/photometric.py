"""CPU restatement (numpy) of the photometric verification at the end of the reference's per-frame path.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / reference arm).

What it restates
  reference call sites: crates/matching-opencv/src/lib.rs:335-389 -- per surviving candidate: warp_affine(frame, M, slide size,
                        WARP_INVERSE_MAP, BORDER_CONSTANT 0) -> to_small_image -> compute_similarity vs slide.small_img; sort by
                        similarity; retain > 0.5; first -> Matching.image.  M is the matrix estimate_affine_partial_2d returns
                        (image_utils.rs:52), i.e. the RANSAC model after 10 Levenberg-Marquardt iterations on the inliers.
  arithmetic          : OpenCV (third party, not under /root/reference):
                        * cv::warpAffine with flags == WARP_INVERSE_MAP: the interpolation bits are 0 = INTER_NEAREST; fixed-point
                          coordinates with AB_BITS = 10: X = (cvRound((M1*y + M2)*1024) + 512 + cvRound(M0*x*1024)) >> 10
                        * cv::LMSolver (calib3d/levmarq.cpp) with AffinePartial2DRefineCallback, maxIters 10, eps FLT_EPSILON
  parity pin          : tests/test_oracle_photometric.py -- warp bit-exact vs cv2.warpAffine; refined matrix within 1e-9 of
                        cv2.estimateAffinePartial2D's (the 4x4 solves are not restated bit for bit: OpenCV uses its Jacobi
                        eigen-solver, this file numpy's LU) -- that is far below what can move a 1/1024-pixel fixed-point coordinate.
"""
from __future__ import annotations

import numpy as np

from . import (MIN_RATING, MIN_RATING_FRACTION, TOP_RATED, TOP_SLIDES, VOTE_RATIO, ransac_affine_partial, similarity, small_size,
               to_small_image)

MIN_SIMILARITY = np.float32(0.5)   # lib.rs:381


def lm_refine(src: np.ndarray, dst: np.ndarray, h0, max_iters: int = 10) -> np.ndarray:
    """cv::LMSolver::run on AffinePartial2DRefineCallback(src, dst): params h = [a, b, tx, ty] of [a -b tx; b a ty]."""
    src = np.asarray(src, np.float64)
    dst = np.asarray(dst, np.float64)

    def compute(h, need_j):
        r = np.empty(2 * len(src))
        r[0::2] = h[0] * src[:, 0] - h[1] * src[:, 1] + h[2] - dst[:, 0]
        r[1::2] = h[1] * src[:, 0] + h[0] * src[:, 1] + h[3] - dst[:, 1]
        j = None
        if need_j:
            j = np.zeros((2 * len(src), 4))
            j[0::2, 0], j[0::2, 1], j[0::2, 2] = src[:, 0], -src[:, 1], 1.0
            j[1::2, 0], j[1::2, 1], j[1::2, 3] = src[:, 1], src[:, 0], 1.0
        return r, j

    eps_d, eps_f = np.finfo(np.float64).eps, float(np.finfo(np.float32).eps)
    x = np.array(h0, np.float64)
    r, j = compute(x, True)
    s = float(r @ r)
    a, v = j.T @ j, j.T @ r
    d_diag = np.diag(a).copy()
    lam, lc, it = 1.0, 0.75, 0
    while True:
        ap = a.copy()
        ap[np.arange(4), np.arange(4)] += lam * d_diag
        d = np.linalg.solve(ap, v)
        xd = x - d
        rd, _ = compute(xd, False)
        sd = float(rd @ rd)
        ds = float(d @ (-(a @ d) + 2 * v))
        rr = (s - sd) / (ds if abs(ds) > eps_d else 1.0)
        if rr > 0.75:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif rr < 0.25:
            t = float(d @ v)
            nu = (sd - s) / (t if abs(t) > eps_d else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0.0:
                maxval = max(eps_d, float(np.abs(np.diag(np.linalg.inv(a))).max()))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if sd < s:
            s, x = sd, xd
            r, j = compute(x, True)
            a, v = j.T @ j, j.T @ r
        it += 1
        if not (it < max_iters and float(np.abs(d).max()) >= eps_f and float(np.abs(r).max()) >= eps_f):
            break
    return x


def refined_matrix(pts_from, pts_to):
    """What estimate_affine_partial_2d(from, to, inliers, RANSAC, 3.0, 2000, 0.99, 10) returns: (M 2x3 or None, mask)."""
    pts_from = np.asarray(pts_from, np.float32).reshape(-1, 2)
    pts_to = np.asarray(pts_to, np.float32).reshape(-1, 2)
    good, mask, model, _ = ransac_affine_partial(pts_from, pts_to)
    if good == 0:
        return None, mask
    h = np.array([model[0], model[3], model[2], model[5]])
    if len(pts_from) > 2:
        m = mask.astype(bool)
        h = lm_refine(pts_from[m], pts_to[m], h)
    return np.array([[h[0], -h[1], h[2]], [h[1], h[0], h[3]]]), mask


def warp_affine_inverse_nearest(src: np.ndarray, m, dw: int, dh: int) -> np.ndarray:
    """cv::warpAffine(src, dst, M, (dw, dh), WARP_INVERSE_MAP, BORDER_CONSTANT, 0): nearest neighbour, 10-bit fixed point."""
    m = np.asarray(m, np.float64).reshape(6)
    x = np.arange(dw, dtype=np.float64)
    adelta = np.rint(m[0] * x * 1024.0).astype(np.int64)
    bdelta = np.rint(m[3] * x * 1024.0).astype(np.int64)
    out = np.zeros((dh, dw) + src.shape[2:], src.dtype)
    h, w = src.shape[:2]
    for y in range(dh):
        x0 = int(np.rint((m[1] * y + m[2]) * 1024.0)) + 512
        y0 = int(np.rint((m[4] * y + m[5]) * 1024.0)) + 512
        xs, ys = (x0 + adelta) >> 10, (y0 + bdelta) >> 10
        ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
        out[y, ok] = src[ys[ok], xs[ok]]
    return out


def decide_frame(idx, dist, page_offsets, frame_pts, pool_pts, frame_bgr, page_grays):
    """The whole tail of match_images_with_frame (lib.rs:268-389) on exact k-NN rows.
    Returns dict(cand, survivors (after the rating gates), rated=[(page, similarity)] sorted by similarity, image=page or -1)."""
    po = np.ascontiguousarray(page_offsets, np.int64)
    npages = len(po) - 1
    by_page = [[] for _ in range(npages)]
    dist = np.asarray(dist, np.float32)
    for q in range(idx.shape[0]):
        if idx[q, 0] < 0:
            continue
        lim = np.float32(dist[q, 0]) * np.float32(VOTE_RATIO)
        for jj in range(idx.shape[1]):
            g = idx[q, jj]
            if g >= 0 and dist[q, jj] < lim:
                by_page[int(np.searchsorted(po, g, side="right") - 1)].append((q, int(g)))
    order = sorted((p for p in range(npages) if by_page[p]), key=lambda p: (-len(by_page[p]), p))[:TOP_SLIDES]
    cand, mats = [], {}
    for p in order:
        mm = by_page[p]
        fr = pool_pts[[g for _, g in mm]]
        to = frame_pts[[q for q, _ in mm]]
        m, mask = refined_matrix(fr, to)
        cand.append((p, len(mm), int(mask.sum())))
        mats[p] = m
    ranked = sorted(cand, key=lambda c: -c[2])[:TOP_RATED]
    best = float(ranked[0][2]) if ranked else 0.0
    surv = [(p, r) for p, _, r in ranked if r > MIN_RATING and (float(r) / best if best else 0.0) > MIN_RATING_FRACTION]
    rated = []
    for p, _ in surv:
        page = page_grays[p]
        ph, pw = page.shape[:2]
        proj = warp_affine_inverse_nearest(frame_bgr, mats[p], pw, ph)              # lib.rs:339-348
        small = to_small_image(proj)                                               # lib.rs:350
        page_small = to_small_image(np.repeat(page[:, :, None], 3, axis=2))        # lib.rs:104-105 (gray replicated to BGR)
        rated.append((p, similarity(small, page_small)))                          # lib.rs:351
    rated.sort(key=lambda t: -t[1])                                                # lib.rs:370 (stable)
    rated = [t for t in rated if t[1] > MIN_SIMILARITY]                            # lib.rs:381
    return dict(cand=cand, survivors=surv, rated=rated, image=rated[0][0] if rated else -1, matrices=mats)
