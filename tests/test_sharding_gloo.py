"""The N>1 host logic on CPU: world_size-2 gloo processes shard frames contiguously, replicate the pool with a
broadcast, match their shard (here with the oracle standing in for the GPU -- tests may use it as the checker) and
rank 0 gathers rows identical to the single-process result."""
import os
import socket

import numpy as np
import pytest

from slideo_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert max(hi - lo for lo, hi in spans) == -(-n // world) or n == 0
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_path):
    import torch.distributed as dist
    import oracle
    import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pages = None
        if rank == 0:
            pages = [synth.hamming_pool(n, seed=400 + i, dup_frac=0.02) for i, n in enumerate((120, 0, 90, 200))]
            desc = np.concatenate(pages)
            offs = np.zeros(len(pages) + 1, np.int32)
            offs[1:] = np.cumsum([len(p) for p in pages])
        else:
            desc = offs = None
        desc, offs = sharding.broadcast_pool_host(desc, offs, src=0)
        page_descs = [desc[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
        lo, hi = sharding.shard_range(n_frames, rank, world)
        local = []
        for f in range(lo, hi):
            q = synth.hamming_queries(desc, 40 + f, seed=500 + f, near_frac=0.8)
            best, votes, _ = oracle.match_frame(q, page_descs)
            local.append((best, votes, len(q)))
        local = np.array(local, np.int32).reshape(-1, 3)
        allr = sharding.gather_results(local, n_frames)
        if rank == 0:
            np.save(out_path, allr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8])
def test_two_rank_run_equals_single_process(tmp_path, n_frames):
    import torch.multiprocessing as mp
    import oracle
    import synth
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, out), nprocs=2, join=True)
    got = np.load(out)
    pages = [synth.hamming_pool(n, seed=400 + i, dup_frac=0.02) for i, n in enumerate((120, 0, 90, 200))]
    desc = np.concatenate(pages)
    want = []
    for f in range(n_frames):
        q = synth.hamming_queries(desc, 40 + f, seed=500 + f, near_frac=0.8)
        best, votes, _ = oracle.match_frame(q, pages)
        want.append((best, votes, len(q)))
    assert np.array_equal(got, np.array(want, np.int32))


def _sift_rows(n, seed):
    """Integer-valued 0..255 rows with norm ~512, like cv2 SIFT descriptors."""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-6) * 512.0
    return np.minimum(np.rint(x), 255).astype(np.float32)


def _sift_frame(pool, offs, f):
    rng = np.random.default_rng(900 + f)
    page = (0, 1, 3)[f % 3]                     # page 2 is empty on purpose
    src = rng.integers(offs[page], offs[page + 1], 25)
    q = np.clip(pool[src] + rng.integers(-2, 3, (25, 128)), 0, 255).astype(np.float32)
    return np.concatenate([q, _sift_rows(10, 950 + f)])


def _sift_match(q, pool, offs):
    import oracle
    idx, dist = oracle.bf_knn_l2(q, pool, 30)
    best, votes, _ = oracle.vote(idx, dist, offs)
    return best, votes, len(q)


def _sift_worker(rank, world, port, n_frames, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        desc = offs = None
        if rank == 0:
            pages = [_sift_rows(n, 600 + i) for i, n in enumerate((150, 80, 0, 120))]
            desc = np.concatenate(pages)
            offs = np.zeros(len(pages) + 1, np.int32)
            offs[1:] = np.cumsum([len(p) for p in pages])
        # SIFT128 pools travel as fp32 rows (n x 128), like slideo_b200_pool_device_view on a SIFT128 ctx
        desc, offs = sharding.broadcast_pool_host(desc, offs, src=0, desc_width=128, dtype=np.float32)
        lo, hi = sharding.shard_range(n_frames, rank, world)
        local = np.array([_sift_match(_sift_frame(desc, offs, f), desc, offs) for f in range(lo, hi)], np.int32).reshape(-1, 3)
        allr = sharding.gather_results(local, n_frames)
        if rank == 0:
            np.save(out_path, allr)
    finally:
        dist.destroy_process_group()


def test_two_rank_sift_pool_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered_sift.npy")
    n_frames = 5
    mp.spawn(_sift_worker, args=(2, _free_port(), n_frames, out), nprocs=2, join=True)
    got = np.load(out)
    pages = [_sift_rows(n, 600 + i) for i, n in enumerate((150, 80, 0, 120))]
    desc = np.concatenate(pages)
    offs = np.zeros(len(pages) + 1, np.int32)
    offs[1:] = np.cumsum([len(p) for p in pages])
    want = np.array([_sift_match(_sift_frame(desc, offs, f), desc, offs) for f in range(n_frames)], np.int32)
    assert np.array_equal(got, want)
    assert list(want[:, 0]) == [0, 1, 3, 0, 1]


def _ag_worker(rank, world, port, out_path):
    import torch.distributed as dist
    import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = (120, 0, 90, 200, 33)                      # 5 pages over 2 ranks: 3 + 2, one of them empty
        lo, hi = sharding.shard_range(len(sizes), rank, world)
        mine = [synth.hamming_pool(n, seed=700 + i, dup_frac=0.0) for i, n in enumerate(sizes)][lo:hi]
        desc = np.concatenate(mine) if mine else np.zeros((0, 32), np.uint8)
        offs = np.zeros(len(mine) + 1, np.int32)
        offs[1:] = np.cumsum([len(p) for p in mine])
        full, offsets = sharding.allgather_pool_host(desc, offs)
        np.savez(out_path % rank, desc=full, offs=offsets)
    finally:
        dist.destroy_process_group()


def test_page_sharded_pool_allgather_world2(tmp_path):
    """The configs[2] pool phase: pages sharded over ranks, pool assembled everywhere by a ragged all-gather."""
    import multiprocessing as mp
    import synth
    port = _free_port()
    out = str(tmp_path / "ag_%d.npz")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_ag_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sizes = (120, 0, 90, 200, 33)
    want = np.concatenate([synth.hamming_pool(n, seed=700 + i, dup_frac=0.0) for i, n in enumerate(sizes)])
    want_offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    for r in range(2):
        got = np.load(out % r)
        assert np.array_equal(got["desc"], want) and np.array_equal(got["offs"], want_offs)
