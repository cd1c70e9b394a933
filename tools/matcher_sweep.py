"""BASELINE.json configs[4]: BFMatcher throughput sweep -- Nq in {1k, 10k, 100k} frame descriptors x Nt in {1k, 10k, 100k, 1M}
pooled slide descriptors, Hamming (K8) and L2 (K10), k = 30, device-resident operands, at 1/2/4/8 GPUs (queries sharded across
ranks, pool replicated: no data-path collective).  Data as SURVEY.md 8(d) states: Hamming rows uniform random bytes (seed 7) with
1 % planted near-duplicates of pool rows among the queries; L2 rows integer-valued 0..255 with row norm ~512 (seed 8).

    python tools/matcher_sweep.py                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/matcher_sweep.py

Rank 0 prints one JSON line per point: time = max over ranks of the device time (CUDA events inside the library),
rates are whole-job (all ranks).  Fractions: K8 (v5, bit-sliced) against the ALU-pipe roofline of its formulation measured on this
GPU by the library's micro-benchmark (LOP3 thread-ops/s / 8.36 LOP3 per pair), K10 against MEASURED_PEAKS.json's sustained bf16
peak with SURVEY.md 8(d)'s algorithmic (2*128 + 3) FLOP per pair."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slideo_b200  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = slideo_b200.Context(slideo_b200.default_config(device=local))
K, REPS = 30, 3
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
bf16_peak = float(peaks.get("bf16_tflops_sustained", 1368.2))
int_peak = ctx.microbench(0) / (1070.0 / 128.0) / 1e9      # Gpair/s per GPU: K8 v5 spends 1070 LOP3 per lane and 128 pooled rows


def hamming_rows(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)


def l2_rows(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty((n, 128), device="cuda").exponential_(1.0, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * 512.0
    return torch.clamp(torch.round(x), max=255).contiguous()


def timed(fn):
    fn()
    ctx.synchronize()
    ctx.timings(reset=True)
    for _ in range(REPS):
        fn()
    ctx.synchronize()
    ms = ctx.timings(reset=True)["ms_knn"] / REPS
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


for nt in (1000, 10000, 100000, 1000000):
    pool_h, pool_l = hamming_rows(nt, 7), l2_rows(nt, 8)
    for nq_total in (1000, 10000, 100000):
        nq = -(-nq_total // world)                       # queries of this rank (the last rank may repeat a few: weak remainder)
        qh, ql = hamming_rows(nq, 70 + rank), l2_rows(nq, 80 + rank)
        n_dup = max(1, nq // 100)                        # 1 % planted near-duplicates: force ties / distance 0..3
        src = torch.randint(0, nt, (n_dup,), device="cuda")
        qh[:n_dup] = pool_h[src]
        qh[: n_dup // 2, 0] ^= 7
        ql[:n_dup] = pool_l[src]
        keys = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        idx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        dist_o = torch.empty((nq, K), dtype=torch.float32, device="cuda")
        ms_h = timed(lambda: ctx.bf_knn_hamming_device(qh.data_ptr(), nq, pool_h.data_ptr(), nt, K, keys.data_ptr()))
        ms_l = timed(lambda: ctx.bf_knn_l2_device(ql.data_ptr(), nq, pool_l.data_ptr(), nt, 128, K, idx.data_ptr(), dist_o.data_ptr()))
        if rank == 0:
            pairs = float(nq) * world * nt
            gp_h, gp_l = pairs / ms_h / 1e6, pairs / ms_l / 1e6
            tf_l = (2.0 * 128.0 + 3.0) * pairs / ms_l / 1e9
            print(json.dumps({"n_gpus": world, "nq": nq * world, "nt": nt, "k": K,
                              "hamming": {"ms": round(ms_h, 4), "gpairs_per_s": round(gp_h, 1), "frac_int_pipe_roofline": round(gp_h / (int_peak * world), 4),
                                          "hbm_bytes_algorithmic": 32 * (nq * world + nt * world) + 8 * K * nq * world},
                              "l2": {"ms": round(ms_l, 4), "gpairs_per_s": round(gp_l, 1), "tflops": round(tf_l, 1),
                                     "frac_bf16_sustained_peak": round(tf_l / (bf16_peak * world), 4)}}), flush=True)
if rank == 0:
    print(json.dumps({"n_gpus": world, "int_pipe_peak_gpairs_per_gpu": int_peak, "bf16_peak_tflops_per_gpu": bf16_peak,
                      "note": "small points are launch/latency bound: one K8/K10 launch per point, no batching across points"}))
if world > 1:
    dist.destroy_process_group()
