"""CPU, world_size 2 over gloo: the multi-process host logic of the frame-sharded path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pixtrack_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = shard.units_of_rank(n_units, world, rank)
        # a stand-in for the tracked result of each unit: deterministic function of the unit id
        T = torch.stack([torch.arange(12, dtype=torch.float32) + 100 * u for u in mine]) if mine else torch.zeros(0, 12)
        failed = torch.tensor([u % 3 == 0 for u in mine], dtype=torch.uint8)
        n_it = torch.tensor([u + 1 for u in mine], dtype=torch.int32)
        table = shard.gather_results(shard.pack_results(mine, T, failed, n_it), n_units)
        slow = shard.max_over_ranks(10.0 + rank, 'cpu')
        out.put((rank, table.clone(), slow))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_units', [7, 8, 1])
def test_units_are_sharded_and_gathered_once_each(n_units):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get() for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, table, slow in got:
        assert table.shape == (n_units, shard.RESULT_WIDTH)
        assert slow == 11.0                                   # max over ranks
        for u in range(n_units):
            assert table[u, 14] == u and table[u, 0] == 100 * u and table[u, 11] == 100 * u + 11
            assert table[u, 12] == float(u % 3 == 0) and table[u, 13] == u + 1
    assert torch.equal(got[0][1], got[1][1])                  # every rank holds the same table


def test_round_robin_assignment():
    assert shard.units_of_rank(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((shard.units_of_rank(10, 4, r) for r in range(4)), [])) == list(range(10))
    assert shard.units_of_rank(2, 4, 3) == []
    with pytest.raises(ValueError):
        shard.units_of_rank(4, 2, 2)
