"""compute-sanitizer target for K10 (knn_l2.cu): a small shape, a shape whose work items run the threshold pre-pass (set
SLIDEO_L2_PRE=4 so that 32 pool tiles are enough), a split shape; results are checked against exact integer arithmetic.
usage: SLIDEO_L2_PRE=4 compute-sanitizer --tool memcheck python tools/sanitize_l2.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200

rng = np.random.default_rng(0)


def rows(n):
    return np.minimum(np.rint(rng.gamma(0.6, 40.0, (n, 128))), 255).astype(np.float32)


with slideo_b200.Context(slideo_b200.default_config()) as c:
    for nq, nt in ((700, 1500), (100 * 128, 8192), (300, 20000)):
        q, t = rows(nq), rows(nt)
        idx, dist = c.bf_knn_l2(q, t, 30)
        for r in (0, nq // 2, nq - 1):
            d2 = ((t.astype(np.int64) - q[r].astype(np.int64)[None, :]) ** 2).sum(-1)
            order = np.lexsort((np.arange(nt), d2))[:30]
            assert np.array_equal(order.astype(np.int32), idx[r]), (nq, nt, r)
            assert np.array_equal(np.sqrt(d2[order].astype(np.float32)), dist[r])
        print(nq, nt, "ok")
print("sanitize target ok")
