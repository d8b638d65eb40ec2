"""world_size-2 check (gloo, CPU) of the multi-GPU recipe: stripes are sharded round-robin across ranks with
no data-path collective; the only communication is the timing reduction bench.py does (MAX over ranks)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, path, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import orc_rust_b200 as ob
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = ob.DecodeJob([path], shard=(rank, world)).plan().stats()
    mine = torch.tensor([st["n_stripes"], st["n_rows"], st["input_bytes"]], dtype=torch.int64)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # stand-in for a per-rank device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put(([v.tolist() for v in allv], t.item()))
    dist.destroy_process_group()


def test_round_robin_sharding_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_orc
    import orc_rust_b200 as ob
    p = str(tmp_path / "li.orc")
    gen_orc.write(gen_orc.lineitem_table(40_000, 2), p, stripe_size=2 << 20)
    total = ob.DecodeJob([p]).plan().stats()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, p, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    parts, tmax = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert sum(x[0] for x in parts) == total["n_stripes"]
    assert sum(x[1] for x in parts) == total["n_rows"]
    assert sum(x[2] for x in parts) == total["input_bytes"]
    assert abs(parts[0][0] - parts[1][0]) <= 1
    assert tmax == 2.0
