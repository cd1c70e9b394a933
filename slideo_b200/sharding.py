"""Multi-GPU plumbing of the path: one process per GPU, frames sharded contiguously, the page-descriptor pool
replicated once.  The reference has no distributed code at all; its only data parallelism is
rayon over frames inside one process (crates/matching-opencv/src/lib.rs:174-221) with the page set shared read-only
(`Arc<Vec<ProcessedImage>>`, lib.rs:134-137) -- the broadcast below is the multi-process equivalent of that Arc.

No collective touches the per-frame data path.  `torch.distributed` is plumbing: NCCL (device pointers of the
library's own pool buffers, zero-copy) on GPUs, gloo (host arrays) in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n units for `rank`: ceil(n / world) per rank (the last ranks may be short/empty)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class _DevView:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_pool_host(desc: Optional[np.ndarray], page_offsets: Optional[np.ndarray], src: int = 0, desc_width: int = 32,
                        dtype=np.uint8):
    """Replicates (desc [n, width], page_offsets [P+1]) from `src` to every rank through host tensors (any backend that
    accepts CPU tensors, i.e. gloo).  Returns the arrays on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    hdr = torch.zeros(2, dtype=torch.int64)
    if rank == src:
        hdr[0], hdr[1] = len(desc), len(page_offsets) - 1
    dist.broadcast(hdr, src)
    n, p = int(hdr[0]), int(hdr[1])
    t_off = torch.from_numpy(np.ascontiguousarray(page_offsets, np.int32)) if rank == src else torch.empty(p + 1, dtype=torch.int32)
    dist.broadcast(t_off, src)
    if rank == src:
        t_desc = torch.from_numpy(np.ascontiguousarray(desc, dtype).reshape(n, desc_width))
    else:
        t_desc = torch.from_numpy(np.empty((n, desc_width), dtype))
    if n:
        dist.broadcast(t_desc, src)
    return t_desc.numpy(), t_off.numpy()


def broadcast_pool_device(ctx, src: int = 0) -> None:
    """ONE NCCL broadcast of the pooled descriptors (ORB256: n x 32 B, SIFT128: n x 128 fp32) straight between the library's device buffers
    (plus the tiny page-offset table and a two-word header).  On return every rank's ctx holds the finalized pool."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    hdr = torch.zeros(5, dtype=torch.int64, device=dev)
    if rank == src:
        n, p = ctx.pool_info()
        _, _, pw, ph = ctx.pool_pages_device_view()
        hdr[0], hdr[1], hdr[2], hdr[3], hdr[4] = n, p, int(ctx.pool_points_device_view()[2]), pw, ph
    dist.broadcast(hdr, src)
    n, p, has_pts, pw, ph = int(hdr[0].item()), int(hdr[1].item()), bool(hdr[2].item()), int(hdr[3].item()), int(hdr[4].item())
    if rank != src:
        ctx.pool_reserve(n, p)
    d_desc, desc_bytes, d_off, off_bytes = ctx.pool_device_view()
    t_off = torch.as_tensor(_DevView(d_off, off_bytes), device=dev)
    dist.broadcast(t_off, src)
    if desc_bytes:
        t_desc = torch.as_tensor(_DevView(d_desc, desc_bytes), device=dev)
        dist.broadcast(t_desc, src)                   # the one payload collective of the whole job
        if has_pts:                                   # keypoint coordinates, only needed by the geometric verification
            d_pt, pt_bytes, _ = ctx.pool_points_device_view()
            dist.broadcast(torch.as_tensor(_DevView(d_pt, pt_bytes), device=dev), src)
            if rank != src:
                ctx.pool_points_device_view(received=True)
    if pw > 0 and ph > 0 and p > 0:                   # the pages' small images, only needed by the warp + similarity gate
        d_sm, sm_bytes, _, _ = ctx.pool_pages_device_view(pw, ph) if rank != src else ctx.pool_pages_device_view()
        dist.broadcast(torch.as_tensor(_DevView(d_sm, sm_bytes), device=dev), src)
    torch.cuda.synchronize()
    if rank != src:
        ctx.pool_commit()


def allgather_pool_host(desc: np.ndarray, page_offsets: np.ndarray, desc_width: int = 32, dtype=np.uint8):
    """Host-tensor variant of `allgather_pool_device` (any backend that accepts CPU tensors, i.e. gloo): every rank passes the
    descriptors / page offsets of its contiguous share of the pages and gets the whole pool back, pages in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    counts_loc = np.diff(np.asarray(page_offsets, np.int64))
    hdr = torch.tensor([len(desc), len(counts_loc)], dtype=torch.int64)
    hdrs = [torch.zeros_like(hdr) for _ in range(world)]
    dist.all_gather(hdrs, hdr)
    ns, ps = [int(h[0]) for h in hdrs], [int(h[1]) for h in hdrs]
    cnt = torch.zeros(max(ps + [1]), dtype=torch.int64)
    cnt[:len(counts_loc)] = torch.from_numpy(counts_loc)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = np.concatenate([c.numpy()[:ps[r]] for r, c in enumerate(cnts)]) if sum(ps) else np.zeros(0, np.int64)
    offsets = np.zeros(len(counts) + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    pad = torch.zeros((max(ns + [1]), desc_width), dtype=torch.from_numpy(np.zeros(1, dtype)).dtype)
    pad[:len(desc)] = torch.from_numpy(np.ascontiguousarray(desc, dtype).reshape(-1, desc_width))
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = np.concatenate([p.numpy()[:ns[r]] for r, p in enumerate(parts)]) if sum(ns) else np.zeros((0, desc_width), dtype)
    return out, offsets


def allgather_pool_device(builder, ctx, dist=None) -> dict:
    """Page-sharded pool build (last row: the escape hatch when rank 0's page extraction would be the serial
    term): every rank extracted a contiguous share of the pages into `builder`; this assembles the whole pool in `ctx` on every
    rank.  The exchange is an all-gather with ragged parts, issued as one in-place NCCL broadcast per source rank straight between
    the library's device buffers (descriptors, and the keypoint coordinates when every rank has them); the page-offset table
    travels as a tiny all-gather of per-page counts.  dist=None: single process, plain device-to-device copies.
    Returns {"name", "bytes", "ms", "n_desc", "n_pages"}."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    n_loc, p_loc = builder.pool_info()
    _, offs_loc = builder.pool_export()
    counts_loc = np.diff(offs_loc.astype(np.int64)).astype(np.int64)
    _, _, has_pts = builder.pool_points_device_view()
    hdr = torch.tensor([n_loc, p_loc, int(has_pts)], dtype=torch.int64, device=dev)
    hdrs = [torch.zeros_like(hdr) for _ in range(world)]
    if dist is not None:
        dist.all_gather(hdrs, hdr)
    else:
        hdrs = [hdr]
    ns = [int(h[0]) for h in hdrs]
    ps = [int(h[1]) for h in hdrs]
    pts = all(bool(int(h[2])) for h in hdrs)
    p_max = max(ps + [1])
    cnt = torch.zeros(p_max, dtype=torch.int64, device=dev)
    cnt[:p_loc] = torch.from_numpy(counts_loc).to(dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    if dist is not None:
        dist.all_gather(cnts, cnt)
    else:
        cnts = [cnt]
    counts = np.concatenate([c.cpu().numpy()[:ps[r]] for r, c in enumerate(cnts)]) if sum(ps) else np.zeros(0, np.int64)
    offsets = np.zeros(len(counts) + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    n_tot, p_tot = int(sum(ns)), int(sum(ps))
    ctx.pool_reserve(n_tot, p_tot)
    d_desc, desc_bytes, d_off, off_bytes = ctx.pool_device_view()
    torch.as_tensor(_DevView(d_off, off_bytes), device=dev).copy_(torch.from_numpy(offsets.view(np.uint8)).to(dev))
    starts = np.concatenate([[0], np.cumsum(ns)])
    width = desc_bytes // max(n_tot, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    moved = 0
    e0.record()

    def exchange(dst_ptr, src_ptr, row_bytes):
        nonlocal moved
        for r in range(world):
            if ns[r] == 0:
                continue
            part = torch.as_tensor(_DevView(dst_ptr + int(starts[r]) * row_bytes, ns[r] * row_bytes), device=dev)
            if r == rank:
                part.copy_(torch.as_tensor(_DevView(src_ptr, ns[r] * row_bytes), device=dev))
            if dist is not None:
                dist.broadcast(part, r)
            moved += ns[r] * row_bytes

    if n_tot:
        s_desc, _, _, _ = builder.pool_device_view()
        exchange(d_desc, s_desc, width)
        if pts:
            s_pt, _, _ = builder.pool_points_device_view()
            d_pt, pt_bytes, _ = ctx.pool_points_device_view()
            exchange(d_pt, s_pt, 8)
            ctx.pool_points_device_view(received=True)
    e1.record()
    torch.cuda.synchronize()
    ctx.pool_commit()
    return {"name": f"all-gather of the page-sharded pool, issued as {world} in-place ncclBroadcast slice(s)" if dist is not None
            else "none (single GPU: device-to-device copy)", "bytes": moved, "ms": e0.elapsed_time(e1), "n_desc": n_tot, "n_pages": p_tot}


def gather_results(local: np.ndarray, n_total: int, world: Optional[int] = None) -> Optional[np.ndarray]:
    """Concatenates the per-rank (best_slide, votes, n_keypoints) rows in rank order on rank 0 (host side, after the
    timed region; results are 12 B/frame)."""
    import torch
    import torch.distributed as dist
    world = world or dist.get_world_size()
    rank = dist.get_rank()
    per = -(-n_total // world)
    buf = np.full((per, 3), -2, np.int32)
    buf[:len(local)] = local
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, outs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy() for o in outs])[:n_total]
