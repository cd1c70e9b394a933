"""Developer probe: run K8 a few times at one geometry (for ncu)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200, synth
nq, nt, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = slideo_b200.Context()
pool = torch.from_numpy(synth.hamming_pool(nt, seed=1, dup_frac=0.0)).cuda()
q = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (nq, 32), dtype=np.uint8)).cuda()
keys = torch.empty((nq, 30), dtype=torch.int32, device="cuda")
for _ in range(reps):
    ctx.bf_knn_hamming_device(q.data_ptr(), nq, pool.data_ptr(), nt, 30, keys.data_ptr())
ctx.synchronize()
t = ctx.timings(reset=True)
print(json.dumps({"nq": nq, "nt": nt, "ms": t["ms_knn"] / reps, "gpairs_per_s": nq * nt / (t["ms_knn"] / reps) / 1e6}))
