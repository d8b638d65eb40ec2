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
    st = ob.DecodeJob(path if isinstance(path, list) else [path], shard=(rank, world)).plan().stats()
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


def test_stripe_assignment_over_several_files(tmp_path):
    """The rule bench.py --gpus N relies on: stripe i of the job's list - counted over all its files in order - goes to
    rank i % N.  Three files with different stripe counts, two ranks: each rank's rows and stored bytes must be exactly
    those of its stripes (not merely add up)."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_orc
    import orc_rust_b200 as ob
    paths = []
    for k, orders in enumerate((9_000, 30_000, 17_000)):
        p = str(tmp_path / f"li{k}.orc")
        gen_orc.write(gen_orc.lineitem_table(orders, 3 + k), p, stripe_size=1 << 20)
        paths.append(p)
    rows = []
    for p in paths:
        fm = ob.ArrowReaderBuilder.try_new(p).file_metadata()
        rows += [fm.stripe_info(i)["number_of_rows"] for i in range(fm.num_stripes)]
    assert len(rows) >= 7 and len(set(rows)) > 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, paths, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    parts, _ = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for r in range(2):
        assert parts[r][0] == len(rows[r::2])
        assert parts[r][1] == sum(rows[r::2])
    # the same rule in one process, for the rank counts the bench runs at
    whole = ob.DecodeJob(paths).plan().stats()
    for world in (4, 8):
        per = [ob.DecodeJob(paths, shard=(r, world)).plan().stats() for r in range(world)]
        assert [s["n_rows"] for s in per] == [sum(rows[r::world]) for r in range(world)]
        assert sum(s["input_bytes"] for s in per) == whole["input_bytes"]
