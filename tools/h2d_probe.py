"""What the host -> device fabric of the box delivers when all ranks copy pinned memory at once (context for bench.py's e2e arm at
N > 1).  torchrun --nproc-per-node N tools/h2d_probe.py"""
import json
import os

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2 << 30
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dst.copy_(src, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
outs = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(outs, t)
else:
    outs = [t]
if rank == 0:
    per = [round(float(o[0]), 1) for o in outs]
    print(json.dumps({"n_gpus": world, "per_gpu_gbs": per, "aggregate_gbs": round(sum(per), 1),
                      "what": "5 x 2 GiB cudaMemcpyAsync pinned host -> device on every rank at the same time"}))
if world > 1:
    dist.destroy_process_group()
