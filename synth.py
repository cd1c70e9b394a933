"""Seeded synthetic slide pages and video frames.  Benchmark / test input only.

Pages : 2001x1125 8-bit gray (pdftocairo's 150 dpi size of the reference fixtures), drawn with cv2 Hershey
        text, shapes, a shared header bar + logo on every page (forces cross-page descriptor ties) and an
        optional noise-texture "figure".
Frames: 1920x1080x3 BGR: page f mod P warped by a small similarity (the reference recommends 1080p screen
        recordings, README.md:11), light sensor noise, a random "speaker" rectangle in a corner;
        5 % of frames are pure clutter (expected: no good match).
Everything is a pure function of (index, P) so every rank / process regenerates identical data.
"""
from __future__ import annotations

import numpy as np

import cv2

PAGE_W, PAGE_H = 2001, 1125
FRAME_W, FRAME_H = 1920, 1080

_WORDS = ("matching slides video frame feature descriptor hamming orb pyramid corner score vote index "
          "lecture theorem proof lemma graph kernel memory bandwidth tensor stream barrier cluster "
          "alpha beta gamma delta epsilon lambda sigma omega result method data model").split()
_FONTS = (cv2.FONT_HERSHEY_SIMPLEX, cv2.FONT_HERSHEY_DUPLEX, cv2.FONT_HERSHEY_COMPLEX, cv2.FONT_HERSHEY_TRIPLEX)


def make_page(p: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + p)
    img = np.full((PAGE_H, PAGE_W), 255, np.uint8)
    # shared header bar + logo glyph (identical on every page)
    cv2.rectangle(img, (0, 0), (PAGE_W - 1, 90), 40, -1)
    cv2.putText(img, "slideo-b200 lecture", (40, 62), cv2.FONT_HERSHEY_DUPLEX, 1.6, 235, 3, cv2.LINE_AA)
    cv2.circle(img, (PAGE_W - 80, 45), 30, 200, 6)
    cv2.line(img, (PAGE_W - 100, 25), (PAGE_W - 60, 65), 200, 5)
    # title
    title = " ".join(rng.choice(_WORDS, 3)) + f" {p}"
    cv2.putText(img, title, (60, 190), _FONTS[int(rng.integers(0, 4))], 2.2, 0, 4, cv2.LINE_AA)
    # body text
    y = 280
    for _ in range(int(rng.integers(6, 15))):
        line = " ".join(rng.choice(_WORDS, int(rng.integers(3, 9))))
        scale = float(rng.uniform(0.9, 1.5))
        cv2.putText(img, line, (int(rng.integers(60, 200)), y), _FONTS[int(rng.integers(0, 4))], scale,
                    int(rng.integers(0, 90)), 2, cv2.LINE_AA)
        y += int(34 * scale + rng.integers(8, 22))
        if y > PAGE_H - 60:
            break
    # shapes
    for _ in range(int(rng.integers(3, 9))):
        kind = int(rng.integers(0, 3))
        x0, y0 = int(rng.integers(1000, PAGE_W - 200)), int(rng.integers(250, PAGE_H - 150))
        col = int(rng.integers(0, 200))
        if kind == 0:
            cv2.rectangle(img, (x0, y0), (x0 + int(rng.integers(40, 260)), y0 + int(rng.integers(30, 140))), col, -1)
        elif kind == 1:
            cv2.circle(img, (x0, y0), int(rng.integers(15, 90)), col, -1)
        else:
            cv2.line(img, (x0, y0), (x0 + int(rng.integers(-200, 200)), y0 + int(rng.integers(-120, 120))), col,
                     int(rng.integers(2, 7)))
    # optional noise-texture figure
    if rng.random() < 0.6:
        fw, fh = int(rng.integers(200, 520)), int(rng.integers(150, 380))
        fx, fy = int(rng.integers(1050, PAGE_W - fw - 20)), int(rng.integers(600, PAGE_H - fh - 20))
        tex = rng.integers(0, 256, (fh // 4 + 1, fw // 4 + 1), dtype=np.uint8)
        tex = cv2.resize(tex, (fw, fh), interpolation=cv2.INTER_CUBIC)
        img[fy:fy + fh, fx:fx + fw] = tex
    return img


_NOISE = None


def _noise_bank() -> np.ndarray:
    global _NOISE
    if _NOISE is None:
        rng = np.random.default_rng(77)
        _NOISE = np.rint(rng.normal(0.0, 2.0, (FRAME_H + 64, FRAME_W + 64, 3))).astype(np.int16)
    return _NOISE


def frame_truth(f: int, npages: int) -> int:
    """Ground-truth page of frame f (-1 for clutter frames)."""
    rng = np.random.default_rng(2_000_000 + f)
    return -1 if rng.random() < 0.05 else f % npages


def make_frame(f: int, npages: int, pages=None) -> np.ndarray:
    """BGR 1080p frame f showing page f mod npages.  `pages` may be a dict/list cache of rendered pages."""
    rng = np.random.default_rng(2_000_000 + f)
    clutter = rng.random() < 0.05
    if clutter:
        small = rng.integers(0, 256, (FRAME_H // 8, FRAME_W // 8), dtype=np.uint8)
        gray = cv2.resize(small, (FRAME_W, FRAME_H), interpolation=cv2.INTER_LINEAR)
        for _ in range(12):
            cv2.putText(gray, " ".join(rng.choice(_WORDS, 4)), (int(rng.integers(0, 1400)), int(rng.integers(60, 1040))),
                        cv2.FONT_HERSHEY_SIMPLEX, float(rng.uniform(1, 3)), int(rng.integers(0, 256)), 3, cv2.LINE_AA)
    else:
        p = f % npages
        page = pages[p] if pages is not None else make_page(p)
        s = FRAME_W / PAGE_W * float(rng.uniform(0.97, 1.03))
        ang = float(rng.uniform(-1.0, 1.0))
        m = cv2.getRotationMatrix2D((PAGE_W / 2, PAGE_H / 2), ang, s)
        m[0, 2] += FRAME_W / 2 - PAGE_W / 2 + float(rng.uniform(-8, 8))
        m[1, 2] += FRAME_H / 2 - PAGE_H / 2 + float(rng.uniform(-8, 8))
        gray = cv2.warpAffine(page, m, (FRAME_W, FRAME_H), flags=cv2.INTER_LINEAR, borderValue=255)
    bgr = cv2.cvtColor(gray, cv2.COLOR_GRAY2BGR)
    oy, ox = int(rng.integers(0, 64)), int(rng.integers(0, 64))
    nb = _noise_bank()[oy:oy + FRAME_H, ox:ox + FRAME_W]
    bgr = np.clip(bgr.astype(np.int16) + nb, 0, 255).astype(np.uint8)
    # "speaker" rectangle in a corner
    sw, sh = 320, 240
    cx = 0 if rng.random() < 0.5 else FRAME_W - sw
    cy = 0 if rng.random() < 0.5 else FRAME_H - sh
    sp = rng.integers(0, 256, (sh // 8, sw // 8, 3), dtype=np.uint8)
    bgr[cy:cy + sh, cx:cx + sw] = cv2.resize(sp, (sw, sh), interpolation=cv2.INTER_CUBIC)
    return bgr


def hamming_pool(n: int, seed: int = 7, dup_frac: float = 0.01) -> np.ndarray:
    """Random 256-bit descriptors with planted duplicates / near-duplicates at distance 0..3 (forces ties)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ndup = int(n * dup_frac)
    if ndup and n > 1:
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        d[dst] = d[src]
        flips = rng.integers(0, 4, ndup)
        for i in range(ndup):
            for _ in range(int(flips[i])):
                b = int(rng.integers(0, 256))
                d[dst[i], b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def hamming_queries(pool: np.ndarray, n: int, seed: int = 9, near_frac: float = 0.5) -> np.ndarray:
    """Queries: half random, half noisy copies of pool rows (so realistic small best distances + exact hits)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    nn = int(n * near_frac)
    if nn and len(pool):
        src = rng.integers(0, len(pool), nn)
        q[:nn] = pool[src]
        nflip = rng.integers(0, 40, nn)
        for i in range(nn):
            bits = rng.integers(0, 256, int(nflip[i]))
            for b in bits:
                q[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return q
