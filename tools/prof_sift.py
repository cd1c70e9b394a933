"""Developer probe: K11 (SIFT) on 1080p frames -- per-frame device time of the extractor, and of the whole SIFT128 frame path
(K11 -> K10 -> vote) against a pool of P synthetic pages.  usage: prof_sift.py [n_frames] [n_pages]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200
import synth

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 16
npg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pages = [synth.make_page(p) for p in range(npg)]
frames = np.stack([synth.make_frame(f, npg, pages) for f in range(nf)])
ctx = slideo_b200.Context(slideo_b200.default_config(descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=int(os.environ.get("SIFT_MAX_BATCH", "8"))))
t0 = time.time()
nk = [ctx.add_page_gray8(p) for p in pages]
ctx.finalize_pool()
t_pages = time.time() - t0
d = torch.from_numpy(frames).cuda()
h, w = frames.shape[1:3]
ctx.match_frames_bgr8_device(d.data_ptr(), nf, w, h)   # warm-up (workspaces grow to the sizes of this run)
ctx.timings(reset=True)
res = ctx.match_frames_bgr8_device(d.data_ptr(), nf, w, h)
tm = ctx.timings(reset=True)
print(json.dumps({"frames": nf, "pages": npg, "pool": int(sum(nk)), "kp_per_frame": float(res[:, 2].mean()),
                  "ms_detect_per_frame": tm["ms_detect"] / nf, "ms_knn_per_frame": tm["ms_knn"] / nf, "ms_total_per_frame": tm["ms_total"] / nf,
                  "frames_per_s": nf / tm["ms_total"] * 1e3, "kernel_launches": tm["kernel_launches"], "s_pages_wall": t_pages,
                  "correct": int(sum(int(res[f, 0]) == f % npg for f in range(nf)))}))
