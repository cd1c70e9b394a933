"""compute-sanitizer target for round 2's new device code: K8 v5 (bit-sliced k-NN: slab boundaries, complemented lists, split tiles, the
radix select, k < 30), the device-driven frame stream (submit / collect over several tickets, a blank frame, the flush of a partial
wave), the rewritten ORB kernels (blur with in-smem reflection on odd sizes, resize strips) and the pool replication views.
    compute-sanitizer --tool memcheck python tools/sanitize_stream.py
    compute-sanitizer --tool racecheck python tools/sanitize_stream.py knn"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200
import synth

what = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(0)


def tex(seed, h, w):
    r = np.random.default_rng(seed)
    small = r.integers(0, 256, (h // 8 + 1, w // 8 + 1)).astype(np.float32)
    big = np.kron(small, np.ones((8, 8), np.float32))[:h, :w]
    big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, (1, 1), (0, 1))) / 4
    return big.astype(np.uint8)


with slideo_b200.Context(slideo_b200.default_config(max_batch=3)) as c:
    if what in ("all", "knn"):
        for nq, nt, k in ((1, 1, 30), (130, 4097, 30), (300, 9000, 7), (5, 20000, 32), (40, 0, 30)):
            pool = synth.hamming_pool(nt, seed=3, dup_frac=0.05) if nt else np.zeros((0, 32), np.uint8)
            q = synth.hamming_queries(pool, nq, seed=4)
            q[::7] |= rng.integers(0, 256, q[::7].shape, dtype=np.uint8)      # dense queries: complemented lists
            idx, dist = c.bf_knn_hamming(q, pool, k)
            assert idx.shape == (nq, k)
    if what in ("all", "frames"):
        pages = [tex(i, 333, 517) for i in range(3)]
        for p in pages:
            c.add_page_gray8(p)
        c.finalize_pool()
        frames = np.stack([np.stack([pages[i % 3]] * 3, axis=2) for i in range(7)])
        frames[3] = 128
        t1 = c.submit_frames_bgr8(np.ascontiguousarray(frames[:4]))
        t2 = c.submit_frames_bgr8(np.ascontiguousarray(frames[4:]))
        r = np.concatenate([c.collect(t1, 4), c.collect(t2, 3)])
        sync = c.match_frames_bgr8(frames)
        print(r.tolist())
        assert np.array_equal(r, sync), (r.tolist(), sync.tolist())
        assert r[3, 2] == 0
        ch, sim = c.mark_changed_bgr8(frames, reset=True)
        assert ch[0]
print("sanitize target ok")
