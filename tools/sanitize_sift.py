"""compute-sanitizer target: K11 on small odd-sized images + one SIFT128 frame-path call (memcheck finds out-of-bounds accesses
that parity tests cannot see).  usage: compute-sanitizer --tool memcheck python tools/sanitize_sift.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200

rng = np.random.default_rng(0)


def tex(h, w, seed):
    r = np.random.default_rng(seed)
    small = r.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    big = np.kron(small, np.ones((8, 8), np.float32))[:h, :w]
    for _ in range(3):
        big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1)) / 3
    return np.clip(big, 0, 255).astype(np.uint8)


with slideo_b200.Context(slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=3, keep_matches=1)) as c:
    for shape in ((67, 121), (120, 203), (33, 47), (9, 8)):
        kf, oc, de = c.extract_sift(tex(*shape, seed=shape[0]))
        print(shape, len(kf))
    pages = [tex(150, 222, 10 + p) for p in range(3)]
    for p in pages:
        c.add_page_gray8(p)
    c.finalize_pool()
    frames = np.stack([np.stack([np.clip(pages[p].astype(np.int16) + rng.integers(-3, 4, pages[p].shape), 0, 255).astype(np.uint8)] * 3, axis=2)
                       for p in (2, 0, 1, 1, 0)])
    print(c.match_frames_bgr8(frames).tolist())
    q = np.minimum(np.rint(rng.gamma(0.6, 40.0, (700, 128))), 255).astype(np.float32)
    t = np.minimum(np.rint(rng.gamma(0.6, 40.0, (1500, 128))), 255).astype(np.float32)
    idx, dist = c.bf_knn_l2(q, t, 30)
    print(idx.shape, float(dist.min()))
