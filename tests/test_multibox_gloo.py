"""world_size-2 gloo test (CPU) of the multi-box dispatcher: box partition, per-box scalar gather,
reductions.  The data path itself needs no collective (independent boxes), so this is all the N>1 host logic."""
import os
import socket

import numpy as np
import pytest

from msmpscu_b200.multibox import MultiBoxDispatcher, partition_boxes


def test_partition_is_contiguous_and_complete():
    for nbox in (1, 2, 7, 512):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                f, c = partition_boxes(nbox, world, r)
                got += list(range(f, f + c))
            assert got == list(range(nbox))
            sizes = [partition_boxes(nbox, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert [partition_boxes(512, 8, r) for r in (0, 7)] == [(0, 64), (448, 64)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nbox, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = MultiBoxDispatcher(nbox)
    # per-box "temperature" and "energy": a function of the global box id, so the gathered table is checkable
    local = np.array([[100.0 + b, -8.9 * b] for b in d.local_boxes()])
    table = d.gather_box_scalars(local)
    tot = d.reduce_sum(np.array([float(d.count)]))
    mx = d.reduce_max(np.array([float(d.first)]))
    q.put((rank, table, tot[0], mx[0], d.first, d.count))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_gather_per_box_scalars():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, nbox = _free_port(), 7
    ps = [ctx.Process(target=_worker, args=(r, 2, port, nbox, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=90) for _ in ps]
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = np.array([[100.0 + b, -8.9 * b] for b in range(nbox)])
    for rank, table, tot, mx, first, count in res:
        assert np.array_equal(table, want)
        assert tot == nbox
        assert mx == 4  # rank 1 starts at box ceil(7/2) = 4
    assert sorted((r[4], r[5]) for r in res) == [(0, 4), (4, 3)]


def test_slab_layers_cover_the_box():
    """host-side plan of the slab decomposition (msmpscu_b200/domain.py): every z-layer of cells has exactly one owner"""
    from msmpscu_b200.domain import slab_layers
    for ncz in (3, 7, 35, 87):
        for world in (1, 2, 3, 4, 8):
            if world > ncz:
                continue
            got = []
            for r in range(world):
                z0, z1 = slab_layers(ncz, world, r)
                assert z1 > z0
                got += list(range(z0, z1))
            assert got == list(range(ncz))
