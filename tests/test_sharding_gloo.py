"""Host-side multi-rank logic on CPU: world_size 2 over gloo (127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fmtuner_sdr_b200 import shard


def test_partitions_are_disjoint_and_complete():
    for world in (1, 2, 4, 8):
        ids = [c for r in range(world) for c in shard.channels_of_rank(r, world, 1250)]
        assert ids == list(range(1250 * world))
        parts = shard.split_total(10_000, world)
        assert sum(len(p) for p in parts) == 10_000
        assert [c for p in parts for c in p] == list(range(10_000))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.channels_of_rank(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.channels_of_rank(rank, world, 5)
    # every rank reports its channels; no sample data is exchanged
    gathered = [None] * world
    dist.all_gather_object(gathered, list(mine))
    ms = shard.max_over_ranks(10.0 + 5.0 * rank)        # slowest rank defines the step time
    dist.barrier()
    q.put((rank, gathered, ms))
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, ms in res:
        assert gathered == [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]]
        assert ms == 15.0
    v = shard.aggregate_throughput(1000, 4, world, 15.0)
    assert abs(v - 2 * 1000 * 4 / 0.015 / 1e6) < 1e-12


def test_single_process_identity():
    assert shard.max_over_ranks(3.5) == 3.5
