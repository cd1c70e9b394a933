"""Developer probe: the ORB frame path on N synthetic 1080p frames x P pages through submit/collect (for ncu launch lists and captures).
usage: python tools/prof_frames.py [frames=128] [pages=50] [steps=3] [max_batch=64]"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth

_P = {}


def _page(p):
    return p, synth.make_page(p)


def _frame(a):
    f, npg = a
    p = f % npg
    if p not in _P:
        _P[p] = synth.make_page(p)
    return f, synth.make_frame(f, npg, _P)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    npg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    mb = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    with mp.get_context("fork").Pool(min(32, os.cpu_count() or 1)) as pool:
        pages = dict(pool.map(_page, range(npg)))
        frames = np.stack([f for _, f in pool.map(_frame, [(i, npg) for i in range(n)], chunksize=2)])
    import torch
    import slideo_b200
    ctx = slideo_b200.Context(slideo_b200.default_config(max_batch=mb))
    for p in range(npg):
        ctx.add_page_gray8(pages[p])
    ctx.finalize_pool()
    dev = torch.from_numpy(frames).cuda()
    ctx.match_frames_bgr8_device(dev.data_ptr(), n, 1920, 1080)
    ctx.timings(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = ctx.submit_frames_bgr8_device(dev.data_ptr(), n, 1920, 1080)
    for _ in range(steps - 1):
        nxt = ctx.submit_frames_bgr8_device(dev.data_ptr(), n, 1920, 1080)
        ctx.collect(prev, n)
        prev = nxt
    res = ctx.collect(prev, n)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    tm = ctx.timings(reset=True)
    print(json.dumps({"frames": n, "pages": npg, "steps": steps, "frames_per_s": n * steps / dt, "ms_total": tm["ms_total"],
                      "ms_knn": tm["ms_knn"], "ms_detect": tm["ms_detect"], "gpairs_per_s": tm["knn_pairs"] / max(tm["ms_knn"], 1e-9) / 1e6,
                      "kp_per_frame": float(res[:, 2].mean())}))


if __name__ == "__main__":
    main()
