"""Developer probe: K10 throughput at one geometry."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200
nq, nt, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = slideo_b200.Context()
g = torch.Generator(device="cuda").manual_seed(1)
def sift(n):
    x = torch.empty((n, 128), device="cuda").exponential_(1.0, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * 512.0
    return torch.clamp(torch.round(x), max=255).contiguous()
t, q = sift(nt), sift(nq)
idx = torch.empty((nq, 30), dtype=torch.int32, device="cuda"); dist = torch.empty((nq, 30), dtype=torch.float32, device="cuda")
for _ in range(2):
    ctx.bf_knn_l2_device(q.data_ptr(), nq, t.data_ptr(), nt, 128, 30, idx.data_ptr(), dist.data_ptr())
ctx.synchronize(); ctx.timings(reset=True)
for _ in range(reps):
    ctx.bf_knn_l2_device(q.data_ptr(), nq, t.data_ptr(), nt, 128, 30, idx.data_ptr(), dist.data_ptr())
tm = ctx.timings(reset=True)
ms = tm["ms_knn"] / reps
print(json.dumps({"nq": nq, "nt": nt, "ms": ms, "gpairs_per_s": nq * nt / ms / 1e6, "tflops": 2 * 144 * nq * nt / ms / 1e9}))
