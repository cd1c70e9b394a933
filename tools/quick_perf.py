"""Developer probe (not the bench): integer-pipe microbenchmarks + K8 throughput at a few geometries."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200
import synth

ctx = slideo_b200.Context()
out = {}
for which, name in ((0, "lop3_ops_per_s"), (1, "popc_ops_per_s"), (2, "mix_pairs_per_s")):
    out[name] = ctx.microbench(which)
print(json.dumps(out))
for nq, nt in ((2048, 100_000), (65536, 100_000), (65536, 1_000_000), (2048, 1_000_000), (1024, 1024), (100_000, 100_000)):
    pool = torch.from_numpy(synth.hamming_pool(nt, seed=1, dup_frac=0.0)).cuda()
    q = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (nq, 32), dtype=np.uint8)).cuda()
    keys = torch.empty((nq, 30), dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.bf_knn_hamming_device(q.data_ptr(), nq, pool.data_ptr(), nt, 30, keys.data_ptr())
    ctx.synchronize()
    ctx.timings(reset=True)
    reps = 5
    for _ in range(reps):
        ctx.bf_knn_hamming_device(q.data_ptr(), nq, pool.data_ptr(), nt, 30, keys.data_ptr())
    t = ctx.timings(reset=True)
    ms = t["ms_knn"] / reps
    print(json.dumps({"nq": nq, "nt": nt, "ms": ms, "gpairs_per_s": nq * nt / ms / 1e6, "launches": t["knn_launches"] / reps}))
