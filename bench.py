#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config: "1080p frames matched/s".

Workload (config.workload): configs[1] = 1k synthetic 1080p BGR frames x 50 rendered slide pages, ORB-2000,
256-bit Hamming brute-force k-NN (k=30) + the reference's 1.05-ratio vote -> (best_slide, votes) per frame.
A "step" is one pass of the hot path over this rank's 1000 frames.  At N>1 every rank gets its own 1000 frames
(frames shard with no data-path collective; the pool is replicated by one NCCL broadcast before the timed region)
-> "scaling": "weak".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = frames/s with the frames already resident in HBM; `e2e` = the same through
slideo_b200_match_frames_bgr8 with HOST (pinned) frames, H2D + D2H inside the timed region.  `roofline` describes the
dominant kernel (K8, brute-force Hamming k-NN): it is integer-pipe bound, not HBM bound,
so the binding fraction is against the measured POPC-pipe ceiling and the HBM fraction is reported beside it.
`cpu_baseline` / `--impl reference` time the reference's CPU path (OpenCV ORB + BFMatcher + vote through cv2, the
library the reference calls) on this box's host cores -- the only place this file touches `oracle/`.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_W, FRAME_H = 1920, 1080
FRAME_BYTES = FRAME_W * FRAME_H * 3
KNN_K = 30

# ----------------------------------------------------------------------------------------------------------------
# worker-process helpers (fork()ed before any CUDA initialisation)
_PAGES = {}
_CPU = {}


def _gen_frame(args):
    f, npages = args
    import synth
    p = f % npages
    if p not in _PAGES:
        _PAGES[p] = synth.make_page(p)
    return f, synth.make_frame(f, npages, _PAGES)


def _gen_page(p):
    import synth
    return p, synth.make_page(p)


def _cpu_init(page_descs):
    import cv2
    cv2.setNumThreads(1)                       # frame-level parallelism like the reference's rayon scope (lib.rs:174-221)
    from oracle import cv2_oracle as co
    _CPU["orb"] = co.make_orb(2000)
    m = cv2.BFMatcher(cv2.NORM_HAMMING)
    m.add([d for d in page_descs if len(d)])
    _CPU["bf"] = m
    _CPU["nonempty"] = [i for i, d in enumerate(page_descs) if len(d)]
    _CPU["npages"] = len(page_descs)
    _CPU["co"] = co


def _cpu_page_desc(p):
    import cv2
    import synth
    from oracle import cv2_oracle as co
    cv2.setNumThreads(1)
    # lib.rs:98,104: imread gray, replicate to 3 channels, ORB converts back
    return p, co.orb_canonical(cv2.cvtColor(synth.make_page(p), cv2.COLOR_GRAY2BGR))[2]


def _cpu_match_frame(args):
    """The reference's per-frame path on the CPU: ORB -> exact BF knn(30) -> vote (lib.rs:264-282)."""
    f, npages = args
    _, frame = _gen_frame((f, npages))
    co = _CPU["co"]
    t0 = time.perf_counter()
    _, _, desc = co.orb_canonical(frame, orb=_CPU["orb"])
    rows = _CPU["bf"].knnMatch(np.ascontiguousarray(desc), KNN_K)
    v = co.vote_rows(rows, len(_CPU["nonempty"]))
    votes = np.zeros(_CPU["npages"], np.int64)
    votes[_CPU["nonempty"]] = v
    best = int(np.argmax(votes)) if votes.max(initial=0) > 0 else -1
    return f, best, int(votes[best]) if best >= 0 else 0, len(desc), time.perf_counter() - t0


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(npages: int, frame_ids, cores: int, steps: int = 1, warmup: int = 0):
    """frames/s of the CPU path on `cores` worker processes; each step matches `frame_ids` once."""
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        page_descs = [None] * npages
        for p, d in pool.imap_unordered(_cpu_page_desc, range(npages)):
            page_descs[p] = d
    times, results = [], None
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(page_descs,)) as pool:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            out = pool.map(_cpu_match_frame, [(f, npages) for f in frame_ids], chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
            results = out
    # (the wall time also contains the synthetic generation of each sampled frame, ~0.3 % of a CPU frame)
    return len(frame_ids) * len(times) / sum(times), times, results, page_descs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "power_w_max": None, "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), power_w_max=max(pw), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = list(range(min(args.frames, max(cores, 8))))
    rate, times, results, _ = cpu_reference_rate(args.pages, sample, cores, steps=args.steps, warmup=args.warmup)
    import cv2
    line = {
        "impl": "reference", "metric": "1080p frames matched/s", "value": rate, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{len(sample)} frames/step of the {args.frames}-frame workload, cv2 {cv2.__version__} "
                                   f"(OpenCV, the library the reference calls; it pins 4.5.2) ORB(2000)+BFMatcher(HAMMING).knnMatch(30)+vote, "
                                   f"one process per core, 1 OpenCV thread each (mirrors rayon per-frame parallelism)"},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"configs[1]: {args.frames} synthetic 1080p BGR frames x {args.pages} slide pages (2001x1125), ORB-2000, "
                        f"256-bit Hamming BF-knn k=30 + 1.05-ratio vote, per GPU",
            "frames_per_gpu": args.frames, "pages": args.pages, "nfeatures": 2000, "knn_k": KNN_K, "max_batch": args.max_batch,
            "parallelism": f"frames sharded over {world} GPU(s), pool replicated (1 NCCL broadcast)" if world > 1 else "1 GPU",
            "l2_policy": f"inputs larger than L2 ({args.frames * FRAME_BYTES / 1e9:.1f} GB of frames per step vs 126 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU per step (configs[1]: 1000)")
    ap.add_argument("--pages", type=int, default=50, help="slide pages in the pool (configs[1]: 50)")
    ap.add_argument("--max-batch", type=int, default=64,
                    help="frames per detection batch (K8 runs on chunks of the pooled query stream, independent of this)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the detail blocks (decision tail, SIFT variant, configs[2])")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    cores = os.cpu_count() or 1
    # ---- CPU baseline first (rank 0, N=1 only): bounded sample of the same workload, on otherwise idle host cores ----
    cpu_baseline = None
    cpu_results = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import cv2
        sample = list(range(min(args.frames, max(cores, 8))))
        rate, times, cpu_results, _ = cpu_reference_rate(args.pages, sample, cores)
        cpu_baseline = {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": f"first {len(sample)} frames of the workload (one pass, {times[0]:.1f} s), cv2 {cv2.__version__} "
                                  f"ORB(2000)+BFMatcher(HAMMING).knnMatch(30)+vote, one process per core, 1 OpenCV thread each"}

    # ---- host-side input generation in fork()ed workers, BEFORE CUDA is initialised in this process ----------
    gen_procs = max(1, min(32, cores // max(world, 1)))
    f_lo = rank * args.frames
    fork = mp.get_context("fork")
    gen_pool = fork.Pool(gen_procs)
    pages_async = gen_pool.map_async(_gen_page, range(args.pages)) if rank == 0 else None
    frames_iter = gen_pool.imap(_gen_frame, [(f_lo + i, args.pages) for i in range(args.frames)], chunksize=4)

    import torch
    import slideo_b200
    from slideo_b200 import sharding

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = slideo_b200.Context(slideo_b200.default_config(device=local_rank, max_batch=args.max_batch))

    # ---- page pool: built on rank 0 by the library's own ORB, replicated by one broadcast --------------------
    t0 = time.perf_counter()
    if rank == 0:
        pages = dict(pages_async.get())
        for p in range(args.pages):
            ctx.add_page_gray8(pages[p])
        ctx.finalize_pool()
    if world > 1:
        sharding.broadcast_pool_device(ctx, src=0)
    pool_n, pool_pages = ctx.pool_info()
    t_pool = time.perf_counter() - t0

    # ---- frames: pinned host copy (e2e arm) + device-resident copy (value arm) --------------------------------
    pin = slideo_b200.PinnedBuffer(args.frames * FRAME_BYTES)
    host = pin.array.reshape(args.frames, FRAME_H, FRAME_W, 3)
    for f, frame in frames_iter:
        host[f - f_lo] = frame
    gen_pool.close()
    gen_pool.join()
    dev = torch.empty(args.frames * FRAME_BYTES, dtype=torch.uint8, device="cuda")
    dev.copy_(torch.from_numpy(pin.array), non_blocking=False)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)

    def timed(submit):
        """K steps, software-pipelined through the library's asynchronous entry points: step s+1 is submitted (uploads, K1-K7
        and K8 launches enqueued) before the results of step s are collected, so the GPU never drains between steps.  Every
        step's work, its H2D copies and the D2H of its results lie inside the bracket."""
        def run(k):
            res, prev = None, submit()
            for _ in range(k - 1):
                nxt = submit()
                res = ctx.collect(prev, args.frames)
                prev = nxt
            return ctx.collect(prev, args.frames)
        res = run(args.warmup)
        ctx.timings(reset=True)
        barrier()
        t0 = time.perf_counter()
        res = run(args.steps)
        ctx.synchronize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = ctx.timings(reset=True)
        barrier()
        # device time of the K steps (CUDA events on the library's streams) and the host bracket; take the larger
        dev_s = tm["ms_total"] * 1e-3
        t = torch.tensor([max(dt, dev_s), dev_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0]), float(t[1]), tm

    # value arm: frames resident in HBM
    sampler.start()
    res_dev, t_dev, t_dev_events, tm_dev = timed(lambda: ctx.submit_frames_bgr8_device(dev.data_ptr(), args.frames, FRAME_W, FRAME_H))
    clocks = sampler.stop()
    # e2e arm: the public host-buffer call, H2D of every frame + D2H of the results inside the timed region
    res_e2e, t_e2e, _, tm_e2e = timed(lambda: ctx.submit_frames_bgr8_ptr(pin.ptr, args.frames, FRAME_W, FRAME_H))
    assert np.array_equal(res_dev, res_e2e), "device-resident and host-buffer arms disagree"
    assert np.array_equal(res_dev, ctx.match_frames_bgr8_ptr(pin.ptr, args.frames, FRAME_W, FRAME_H)), "submit/collect and the synchronous call disagree"

    total_frames = args.frames * world * args.steps
    value = total_frames / t_dev
    e2e = total_frames / t_e2e

    # ---- roofline of the dominant kernel (K8) -----------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    popc_rate = ctx.microbench(1)       # thread-level POPC/s over the whole GPU, measured now on this device
    lop3_rate = ctx.microbench(0)
    mix_rate = ctx.microbench(2)
    knn_s = tm_dev["ms_knn"] * 1e-3
    launches = max(int(tm_dev["knn_launches"]), 1)
    pairs = float(tm_dev["knn_pairs"])
    nq_total = pairs / max(pool_n, 1)
    alg_bytes = 32.0 * (nq_total + pool_n * launches) + 8.0 * KNN_K * nq_total     # 32(Nq+Nt) + 8 k Nq per launch
    ach_pairs = pairs / knn_s / 1e9
    # two-pipe integer roofline of the carry-save formulation: 4 POPC per pair on the XU pipe, 13 LOP3 per
    # pair on the ALU pipe; both rates measured now on this GPU by the library's own micro-benchmarks
    peak_pairs = min(popc_rate / 4.0, lop3_rate / 13.0) / 1e9
    roofline = {
        "bound": "int-pipe", "kernel": "knn_hamming_kernel (K8, + split merge)", "achieved": ach_pairs, "peak": peak_pairs,
        "unit": "Gpair/s", "frac": ach_pairs / peak_pairs,
        "peak_source": "measured now on this GPU: min(POPC thread-ops/s / 4 POPC per pair, LOP3 thread-ops/s / 13 LOP3 per pair); "
                       "K8 is integer-pipe bound, not HBM- or tensor-bound (SURVEY.md 8d); SURVEY's naive 8-POPC ceiling would be popc/8",
        "naive_popc8_ceiling_gpairs": popc_rate / 8.0 / 1e9,
        "avg_launch_ms": 1e3 * knn_s / launches, "pairs_per_launch": pairs / launches, "share_of_step": knn_s / max(t_dev_events, 1e-9),
        "lop3_ops_per_s": lop3_rate, "popc_ops_per_s": popc_rate, "mix_ceiling_gpairs": mix_rate / 1e9,
        "hbm": {"bound": "hbm", "achieved": alg_bytes / knn_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / knn_s / 1e9 / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
        "traffic": None,
    }
    prof = os.path.join(ROOT, "profiles", "k8_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- extra (rank 0, N=1): the same step with the reference's complete decision tail switched on (SURVEY 8(f) ranks 1-2:
    #      RANSAC rating gate lib.rs:284-333 + warp/similarity gate lib.rs:335-389) ----
    verify_detail = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            vctx = slideo_b200.Context(slideo_b200.default_config(device=local_rank, max_batch=args.max_batch, geometric_verification=2))
            for p in range(args.pages):
                vctx.add_page_gray8(pages[p])
            vctx.finalize_pool()
            vctx.match_frames_bgr8_device(dev.data_ptr(), args.frames, FRAME_W, FRAME_H)
            vctx.timings(reset=True)
            t0 = time.perf_counter()
            rv = vctx.match_frames_bgr8_device(dev.data_ptr(), args.frames, FRAME_W, FRAME_H)
            dtv = time.perf_counter() - t0
            tmv = vctx.timings(reset=True)
            dec = vctx.get_decisions(0, args.frames)
            import synth as _synth
            good = 0
            for i in range(args.frames):
                tr = _synth.frame_truth(f_lo + i, args.pages)
                good += int(dec[i]["image"] == (tr if tr >= 0 else -1))
            verify_detail = {"frames_per_s": args.frames / dtv, "ms_verify_per_step": tmv["ms_verify"], "ms_per_step": 1e3 * dtv,
                             "final_decisions_matching_ground_truth": good, "frames": args.frames,
                             "same_vote_results": bool(np.array_equal(rv, res_dev))}
            vctx.close()
        except Exception as ex:  # the extra must never break the contract line
            verify_detail = {"error": str(ex)}

    # ---- extra (rank 0, N=1): the SIFT-128 / L2 variant of the same path (north_star; BASELINE configs[3] geometry at this run's
    #      page count): K11 SIFT on the GPU for pages and frames -> K10 tcgen05 L2 k-NN -> vote, on a bounded sample of the frames ----
    sift_detail = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            ns = min(args.frames, 64)
            sctx = slideo_b200.Context(slideo_b200.default_config(device=local_rank, max_batch=16, descriptor_kind=slideo_b200.ffi.DESC_SIFT128))
            t0 = time.perf_counter()
            for p in range(args.pages):
                sctx.add_page_gray8(pages[p])
            sctx.finalize_pool()
            t_spool = time.perf_counter() - t0
            s_n, _ = sctx.pool_info()
            sctx.match_frames_bgr8_device(dev.data_ptr(), ns, FRAME_W, FRAME_H)   # warm-up: workspaces grow to this sample's sizes
            sctx.timings(reset=True)
            s_sampler = ClockSampler(local_rank)
            s_sampler.start()
            t0 = time.perf_counter()
            for _ in range(3):
                rs = sctx.match_frames_bgr8_device(dev.data_ptr(), ns, FRAME_W, FRAME_H)
            dts = (time.perf_counter() - t0) / 3
            s_clocks = s_sampler.stop()
            tms = sctx.timings(reset=True)
            for key in ("ms_detect", "ms_knn", "knn_pairs"):
                tms[key] = tms[key] / 3
            bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1368.2)))
            k10_tflops = 2.0 * 144.0 * float(tms["knn_pairs"]) / max(tms["ms_knn"], 1e-9) / 1e9
            sift_detail = {"frames": ns, "frames_per_s": ns / dts, "pool_descriptors": s_n, "pool_build_s": t_spool,
                           "keypoints_per_frame": float(np.mean(rs[:, 2])), "ms_k11_per_frame": tms["ms_detect"] / ns,
                           "ms_k10_per_frame": tms["ms_knn"] / ns,
                           "k10_roofline": {"bound": "tensor", "achieved": k10_tflops, "peak": bf16_peak, "unit": "TFLOP/s",
                                            "frac": k10_tflops / bf16_peak, "flops": "2*(128+16)*Nq*Nt per launch"},
                           "frames_with_truth_match": int(sum(1 for i in range(ns) if _truth_ok(rs, f_lo + i, i, args.pages))),
                           "clocks": {"sm_mhz": s_clocks["sm_mhz"], "reasons": s_clocks["reasons"], "power_w_max": s_clocks["power_w_max"]}}
            sctx.close()
        except Exception as ex:  # the extra must never break the contract line
            sift_detail = {"error": str(ex)}

    if rank == 0:
        parity = None
        if cpu_results is not None:
            # identical best-slide assignments to the CPU path on the sampled frames (bench-side sanity, not the parity suite)
            parity = all(tuple(res_dev[f][:3]) == (b, v, n) for f, b, v, n, _ in cpu_results)
        line = {
            "metric": "1080p frames matched/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": args.frames * FRAME_BYTES,
                    "d2h_bytes_per_step": args.frames * 12, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(tm_dev["kernel_launches"]),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "power_w_max": clocks["power_w_max"], "samples": clocks["samples"]},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "detail": {"device_ms_per_step_events": 1e3 * t_dev_events / args.steps, "ms_detect_per_step": tm_dev["ms_detect"] / args.steps,
                       "ms_knn_per_step": tm_dev["ms_knn"] / args.steps, "pool_descriptors": pool_n, "pool_pages": pool_pages,
                       "pool_build_s": t_pool, "keypoints_per_frame": float(np.mean(res_dev[:, 2])),
                       "frames_with_truth_match": int(sum(1 for i in range(args.frames) if _truth_ok(res_dev, f_lo + i, i, args.pages))),
                       "cpu_sample_matches_gpu": parity, "descriptor_pairs_per_s": pairs * world / t_dev,
                       "with_geometric_verification": verify_detail, "sift128_variant": sift_detail},
        }
        print(json.dumps(line), flush=True)
    pin.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _truth_ok(res, f_global, i, npages):
    import synth
    t = synth.frame_truth(f_global, npages)
    return t < 0 or res[i, 0] == t


if __name__ == "__main__":
    main()
