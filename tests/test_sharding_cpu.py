"""Host-side multi-GPU logic on the CPU: world_size-2/3 gloo process groups exercise the shard
arithmetic and the pose-record gather of posenet/sharding.py (the only collective of the path).
The records are synthetic: the gather moves bytes, it never computes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from posenet import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(n, P, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((n, P), generator=g, dtype=torch.float64), torch.rand((n, P, 17), generator=g, dtype=torch.float64),
            torch.rand((n, P, 17, 2), generator=g, dtype=torch.float64) * 513, torch.randn((n, P, 17, 2), generator=g, dtype=torch.float64))


def _worker(rank, world, port, n_total, P, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _records(n_total, P, seed=7)
        b, e = sharding.shard_bounds(n_total, rank, world)
        got = sharding.gather_pose_records(*[t[b:e] for t in full], n_total=n_total)
        got2 = sharding.gather_pose_records(*[t[b:e] for t in full])           # n_total discovered by all-reduce
        ok = all(torch.equal(a, f) for a, f in zip(got, full)) and all(torch.equal(a, f) for a, f in zip(got2, full))
        np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([int(ok)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 8), (2, 7), (3, 10)])
def test_gather_pose_records_gloo(tmp_path, world, n_total):
    mp.spawn(_worker, args=(world, _free_port(), n_total, 5, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / ("ok%d.npy" % r))[0] == 1, "rank %d gathered wrong records" % r


@pytest.mark.parametrize("n,world", [(64, 8), (7, 2), (10, 3), (3, 4), (0, 2)])
def test_shard_bounds_partition(n, world):
    blocks = [sharding.shard_bounds(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in blocks]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_pack_unpack_roundtrip_and_identity_without_group():
    rec = _records(4, 10, seed=1)
    rows = sharding.pack_pose_records(*rec)
    assert rows.shape == (4, 10 * 86)
    assert all(torch.equal(a, b) for a, b in zip(sharding.unpack_pose_records(rows, 10), rec))
    assert all(torch.equal(a, b) for a, b in zip(sharding.gather_pose_records(*rec), rec))
