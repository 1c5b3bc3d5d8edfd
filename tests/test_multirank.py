"""world_size-2 gloo test (CPU) of the multi-GPU path: restart sharding by rank and the single all_gather of
[loss | flat parameters] after which every rank selects the same winner."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, restarts, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pilco_b200.policy_opt import shard_restarts, gather_table, select_best
    rng = np.random.RandomState(0)                      # same table on every rank; each fills only its rows
    full = rng.rand(restarts, 1 + P)
    full[3, 0] = full[5, 0] = -1.0                      # tie between restarts 3 and 5 -> 3 must win
    local = np.full_like(full, np.nan)
    mine = shard_restarts(restarts, rank, world)
    local[mine] = full[mine]
    out = gather_table(local, restarts, rank, world, dist)
    q.put((rank, np.allclose(out, full), select_best(out)))
    dist.destroy_process_group()


def test_two_rank_gather_and_select():
    world, restarts, P = 2, 7, 5
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, restarts, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [3, 3]
