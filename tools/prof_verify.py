"""Developer probe: the full decision tail (cfg.geometric_verification = 2: K12 RANSAC gate + K14 warp / similarity gate) on N
synthetic 1080p frames x P pages through match_frames (for ncu launch lists).  usage: python tools/prof_verify.py [frames=256] [pages=50]"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from prof_frames import _frame, _page  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    npg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    with mp.get_context("fork").Pool(min(32, os.cpu_count() or 1)) as pool:
        pages = dict(pool.map(_page, range(npg)))
        frames = np.stack([f for _, f in pool.map(_frame, [(i, npg) for i in range(n)], chunksize=2)])
    import torch
    import slideo_b200
    ctx = slideo_b200.Context(slideo_b200.default_config(geometric_verification=2))
    for p in range(npg):
        ctx.add_page_gray8(pages[p])
    ctx.finalize_pool()
    dev = torch.from_numpy(frames).cuda()
    ctx.match_frames_bgr8_device(dev.data_ptr(), n, 1920, 1080)
    ctx.timings(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.match_frames_bgr8_device(dev.data_ptr(), n, 1920, 1080)
    dt = time.perf_counter() - t0
    tm = ctx.timings(reset=True)
    dec = ctx.get_decisions(0, n)
    ok = sum(int(d["image"] == (synth.frame_truth(f, npg) if synth.frame_truth(f, npg) >= 0 else -1)) for f, d in enumerate(dec))
    print(json.dumps({"frames": n, "pages": npg, "frames_per_s": n / dt, "ms_total": tm["ms_total"], "ms_knn": tm["ms_knn"],
                      "ms_detect": tm["ms_detect"], "ms_verify": tm.get("ms_verify"), "decisions_ok": ok}))


if __name__ == "__main__":
    main()
