"""K10 (tcgen05 L2 k-NN) probe: times slideo_b200_bf_knn_l2_device on a few shapes and prints TFLOP/s on the algorithmic
2*128 + 3 FLOP per pair, the fraction of the sustained bf16 peak, and a checksum of (idx, dist) so that two builds / knob settings
can be compared for identical results.  Developer knobs of knn_l2.cu are read from the environment by the library
(SLIDEO_L2_PROF: per-phase cycle counters on stderr, SLIDEO_L2_DEBUG, SLIDEO_L2_TRIGGER, SLIDEO_L2_NO_SPLIT).

    python tools/k10_probe.py [nq:nt ...]          default: 65536:1000000 136000:347000 100000:100000
"""
import json
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slideo_b200  # noqa: E402

torch.cuda.set_device(0)
ctx = slideo_b200.Context(slideo_b200.default_config(device=0))
K, REPS = 30, 3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1380.1))
except Exception:
    peak = 1380.1


def l2_rows(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty((n, 128), device="cuda").exponential_(1.0, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * 512.0
    return torch.clamp(torch.round(x), max=255).contiguous()


shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(65536, 1000000), (136000, 347000), (100000, 100000)]
for nq, nt in shapes:
    q, t = l2_rows(nq, 80), l2_rows(nt, 8)
    idx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    run = lambda: ctx.bf_knn_l2_device(q.data_ptr(), nq, t.data_ptr(), nt, 128, K, idx.data_ptr(), dist.data_ptr())  # noqa: E731
    run()
    ctx.synchronize()
    ctx.timings(reset=True)
    for _ in range(REPS):
        run()
    ctx.synchronize()
    ms = ctx.timings(reset=True)["ms_knn"] / REPS
    tf = nq * nt * (2 * 128 + 3) / (ms * 1e-3) / 1e12
    crc = zlib.crc32(idx.cpu().numpy().tobytes()) ^ zlib.crc32(dist.cpu().numpy().tobytes())
    print(json.dumps({"nq": nq, "nt": nt, "ms": round(ms, 3), "tflops_algorithmic": round(tf, 1), "frac_sustained": round(tf / peak, 4),
                      "crc": f"{crc:08x}", "env": {k: v for k, v in os.environ.items() if k.startswith("SLIDEO_L2")}}), flush=True)
