"""Host-side multi-GPU logic on CPU: view sharding and the final gather over a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdf_net_b200 import sharding


def test_shard_units_partition():
    for n in (0, 1, 7, 49, 49 * 22):
        for world in (1, 2, 4, 8):
            shards = [sharding.shard_units(n, r, world) for r in range(world)]
            assert sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        sharding.shard_units(4, 2, 2)
    assert sharding.units_of_scans([2, 3]) == [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_units, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_units(num_units, rank, world)
        # a unit's "map" encodes its global index, so the gathered order can be checked exactly
        local = torch.stack([torch.full((2, 3), float(i)) for i in mine]) if mine else torch.zeros((0, 2, 3))
        out = sharding.gather_maps(local, num_units)
        if rank == 0:
            ret["ok"] = bool(torch.equal(out[:, 0, 0], torch.arange(num_units, dtype=torch.float32))) and out.shape == (num_units, 2, 3)
        else:
            assert out is None
        # per-rank throughput bookkeeping of bench.py: max over ranks of the elapsed time
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_units", [7, 8])
def test_gather_maps_world_size_2(num_units):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), num_units, ret), nprocs=world, join=True)
        assert ret.get("ok") is True
