"""World-size-2 gloo test of the N>1 host logic (window -> rank assignment, max-over-ranks timing,
summary gather).  The GPU data path has no collective."""
import os
import torch.multiprocessing as mp
import torch.distributed as dist


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvio_b200 import sharding
    n = 8
    mine = sharding.shard_windows(n, world, rank)
    local = {i: {"window": i, "rank": rank, "iterations": 3 + i} for i in mine}
    allsum = sharding.gather_summaries(local, n)
    t = sharding.max_over_ranks(1.0 + rank)
    dist.barrier()
    q.put((rank, mine, [s["rank"] for s in allsum], t))
    dist.destroy_process_group()


def test_two_rank_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29531 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    assert res[0][2] == res[1][2] == [0, 1, 0, 1, 0, 1, 0, 1]
    assert res[0][3] == res[1][3] == 2.0
