"""N-rank check (torchrun): pool built on rank 0, replicated by the NCCL broadcast (descriptors + keypoint coordinates + the pages'
small images); every rank matches + verifies (both gates) the SAME frames; results must be byte-identical across ranks and equal to
rank 0's own.  Then the page-sharded build (configs[2] pool phase): every rank extracts a share of the pages, the pool is assembled
everywhere by the ragged all-gather and must equal the pool rank 0 built alone."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slideo_b200, synth
from slideo_b200 import sharding

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NP, NF = 6, 8
ctx = slideo_b200.Context(slideo_b200.default_config(device=local, geometric_verification=2, keep_matches=1))
pages = [synth.make_page(p) for p in range(NP)]
if rank == 0:
    for p in pages:
        ctx.add_page_gray8(p)
    ctx.finalize_pool()
sharding.broadcast_pool_device(ctx, src=0)
frames = np.stack([synth.make_frame(f, NP, pages) for f in range(NF)])
res = ctx.match_frames_bgr8(frames)
ver = ctx.get_verification(0, NF)
dec = ctx.get_decisions(0, NF)
flat = np.array([[len(v["cand"]), len(v["survivors"]), d["image"], int(np.float32(d["rated"][0][1]).view(np.int32)) if d["rated"] else 0]
                 + [x for c in v["cand"] for x in c] + [0] * (3 * (40 - len(v["cand"]))) for v, d in zip(ver, dec)], np.int32)
t = torch.from_numpy(np.concatenate([res.reshape(NF, -1), flat], axis=1)).cuda()
outs = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(outs, t)
if rank == 0:
    same = all(torch.equal(outs[0], o) for o in outs)
    print("ranks agree:", same, "| frame results:", res[:, :2].tolist(), "| survivors:", [v["survivors"] for v in ver],
          "| decisions:", [d["image"] for d in dec])
    assert same
# page-sharded pool build + ragged all-gather == the pool rank 0 built alone
builder = slideo_b200.Context(slideo_b200.default_config(device=local))
lo, hi = sharding.shard_range(NP, rank, world)
for p in pages[lo:hi]:
    builder.add_page_gray8(p)
builder.finalize_pool()
ctx3 = slideo_b200.Context(slideo_b200.default_config(device=local))
coll = sharding.allgather_pool_device(builder, ctx3, dist)
res3 = ctx3.match_frames_bgr8(frames)
ok3 = torch.tensor([int(np.array_equal(res3, res))], device="cuda")
dist.all_reduce(ok3, op=dist.ReduceOp.MIN)
if rank == 0:
    print("page-sharded pool == rank-0 pool on every rank:", bool(ok3.item()), "|", coll)
    assert ok3.item() == 1 and ctx3.pool_info() == ctx.pool_info()
# SIFT128 variant: pool built by K11 on rank 0, fp32 rows broadcast, every rank derives its own bf16 operands
ctx2 = slideo_b200.Context(slideo_b200.default_config(device=local, descriptor_kind=slideo_b200.ffi.DESC_SIFT128, max_batch=4))
small = [np.ascontiguousarray(p[60:700, 100:1060]) for p in pages[:3]]
if rank == 0:
    for p in small:
        ctx2.add_page_gray8(p)
    ctx2.finalize_pool()
sharding.broadcast_pool_device(ctx2, src=0)
rng = np.random.default_rng(5)
fr = np.stack([np.stack([np.clip(small[p].astype(np.int16) + rng.integers(-3, 4, small[p].shape), 0, 255).astype(np.uint8)] * 3, axis=2) for p in (2, 0, 1, 2)])
res2 = ctx2.match_frames_bgr8(fr)
t2 = torch.from_numpy(res2.astype(np.int32)).cuda()
outs2 = [torch.empty_like(t2) for _ in range(world)]
dist.all_gather(outs2, t2)
if rank == 0:
    same2 = all(torch.equal(outs2[0], o) for o in outs2)
    print("SIFT ranks agree:", same2, "| frame results:", res2.tolist())
    assert same2 and res2[:, 0].tolist() == [2, 0, 1, 2]
dist.destroy_process_group()
