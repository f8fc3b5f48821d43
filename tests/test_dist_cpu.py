"""N > 1 host logic on CPU: world_size-2 gloo run of the shard-by-recording plan + host-side gather."""

import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from orcai_b200.sharding import assign_rows, gather_to_rank0, rows_for_rank


def test_lpt_plan_properties():
    costs = [5, 1, 9, 3, 3, 7, 2, 8, 4, 6, 100]
    for world in (1, 2, 4, 8):
        plan = assign_rows(costs, world)
        assert sorted(i for p in plan for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in plan]
        assert max(loads) >= 100 and (world == 1 or max(loads) <= 100 + 9)
    assert assign_rows([], 4) == [[], [], [], []]
    assert assign_rows([1, 1, 1, 1], 2) == [[0, 2], [1, 3]]  # deterministic tie-breaking


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [float(10 + (i * 7) % 13) for i in range(25)]
    mine = rows_for_rank(costs, world, rank)
    # stand-in for per-recording results: (row, number of label rows) -- only the gather is exercised
    results = [(i, i % 5) for i in mine]
    got = gather_to_rank0(results, world, rank)
    dist.barrier()
    if rank == 0:
        q.put(got)
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == 2
    rows = sorted(i for part in got for i, _ in part)
    assert rows == list(range(25))  # disjoint cover: every recording annotated exactly once
    assert all(k == i % 5 for part in got for i, k in part)


def test_rank_shard_from_launcher_environment(monkeypatch):
    """The table run of one worker process: RANK / WORLD_SIZE (torchrun) or ORCAI_B200_SHARD name its share; all workers derive
    the same plan from the file sizes, so the shares are a disjoint cover without any exchange."""
    from orcai_b200 import predict

    for k in ("RANK", "WORLD_SIZE", "ORCAI_B200_SHARD"):
        monkeypatch.delenv(k, raising=False)
    assert predict._rank_shard() is None
    monkeypatch.setenv("WORLD_SIZE", "1"); monkeypatch.setenv("RANK", "0")
    assert predict._rank_shard() is None                       # a single process is not a shard
    monkeypatch.setenv("WORLD_SIZE", "8"); monkeypatch.setenv("RANK", "5")
    assert predict._rank_shard() == (5, 8)
    monkeypatch.setenv("ORCAI_B200_SHARD", "1/4")               # explicit setting wins over the launcher's
    assert predict._rank_shard() == (1, 4)
    costs = [float(3 + (i * 5) % 11) for i in range(40)]
    shares = [rows_for_rank(costs, 8, r) for r in range(8)]
    assert sorted(i for sh in shares for i in sh) == list(range(40))
    loads = [sum(costs[i] for i in sh) for sh in shares]
    assert max(loads) - min(loads) <= max(costs)
