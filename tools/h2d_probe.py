"""Pinned host -> device copy bandwidth with every rank copying at once (torchrun): the platform's ceiling for the
end-to-end figure, measured with plain torch copies (none of this repo's code).

    python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8).pin_memory()
src.fill_(7)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank < active:
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([4 * n / dt / 1e9 if rank < active else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    out[f"{active}_ranks_copying"] = {"aggregate_GBps": round(t.item(), 1), "per_rank_GBps": round(t.item() / active, 1)}
if rank == 0:
    print(json.dumps({"h2d_pinned_copy": out, "world": world}))
if world > 1:
    dist.destroy_process_group()
