"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (satellite_approximation_b200/multi.py) --
work assignment without communication, the id broadcast that sa_dist_init needs, bench.py's max-over-ranks timing rule,
and the row partition + halo protocol of the row-decomposed solve (csrc/dist.cu) replayed on numpy arrays: the 5-point
operator applied slice by slice with exchanged halo rows must equal the operator applied to the whole grid."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from satellite_approximation_b200 import multi


def test_pack_regions_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    sizes = np.exp(rng.uniform(np.log(100), np.log(5e4), 10000)).astype(np.int64)  # BASELINE.json configs[3] areas
    for world in (1, 2, 4, 8):
        bins = multi.pack_regions(sizes, world)
        flat = sorted(i for b in bins for i in b)
        assert flat == list(range(len(sizes)))
        load = [int(sizes[b].sum()) for b in bins]
        assert max(load) - min(load) <= int(sizes.max())  # greedy largest-first: within one item of even
    assert multi.pack_regions([], 3) == [[], [], []]
    assert multi.pack_regions([5], 2) == [[0], []]


def test_region_shard_masks_partition_the_mask():
    from scipy.ndimage import label

    from satellite_approximation_b200 import synth

    mask = synth.region_mask(300, 340, 30, area_lo=30.0, area_hi=3000.0, seed=4)
    lab, k = label(mask)
    for world in (1, 2, 3, 8):
        shards = [multi.region_shard_mask(lab, k, world, r) for r in range(world)]
        total = np.zeros(mask.shape, int)
        for m, labels in shards:
            total += m
            assert set(np.unique(lab[m])) == set(labels)
        assert np.array_equal(total, mask.astype(int))  # every invalid pixel in exactly one shard
        assert sorted(l for _, ls in shards for l in ls) == list(range(1, k + 1))


def test_round_robin_covers_every_item_once():
    for n, world in ((13, 8), (13, 2), (3, 8), (0, 4)):
        got = sorted(i for r in range(world) for i in multi.round_robin(n, world, r))
        assert got == list(range(n))


def test_row_partition_alignment_and_cover():
    for rows in (1, 31, 32, 700, 10980, 20000):
        for world in (1, 2, 4, 8):
            bounds, levels = multi.row_partition(rows, world)
            block = 32 << (levels - 1)
            assert bounds[0] == 0 and bounds[-1] >= rows and len(bounds) == world + 1
            assert all(b % block == 0 for b in bounds) and all(a <= b for a, b in zip(bounds, bounds[1:]))
            sizes = [b - a for a, b in zip(bounds, bounds[1:])]
            assert max(sizes) - min(sizes) <= block  # as even as whole blocks allow
            assert all(a >= b for a, b in zip(sizes, sizes[1:]))  # empty slices only at the tail (dist_halo relies on it)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _apply(x):
    p = np.pad(x, 1)
    return 4.0 * x - (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id travels from rank 0 to everyone
        payload = bytes(range(128))
        got = multi.broadcast_bytes(payload if rank == 0 else None, 128, 0)
        assert got == payload
        # 2. timing rule: max over ranks; independent scenes add up, one shared system is counted once
        ms, units = multi.reduce_step(10.0 + rank, 100.0 * (rank + 1), one_system=False)
        assert ms == 10.0 + world - 1 and units == 100.0 * world * (world + 1) / 2
        ms, units = multi.reduce_step(10.0 + rank, 7.0, one_system=True)
        assert ms == 10.0 + world - 1 and units == 7.0
        # 3. every rank derives the same assignment of independent work without talking
        sizes = [int(v) for v in np.random.default_rng(3).integers(100, 50000, 200)]
        mine = multi.pack_regions(sizes, world)[rank]
        allb = [None] * world
        dist.all_gather_object(allb, mine)
        assert sorted(i for b in allb for i in b) == list(range(200))
        # 4. row decomposition: slice + one halo row from each neighbour == the whole-grid operator
        rows, cols = 300, 70
        x = np.random.default_rng(5).standard_normal((rows, cols))
        bounds, _ = multi.row_partition(rows, world, levels=1)
        lo, hi = min(bounds[rank], rows), min(bounds[rank + 1], rows)
        local = np.zeros((hi - lo + 2, cols))
        local[1:-1] = x[lo:hi]
        up, down = rank - 1, rank + 1
        reqs = []
        if up >= 0 and hi > lo:
            reqs.append(dist.isend(torch.from_numpy(local[1].copy()), up))
            top = torch.zeros(cols, dtype=torch.float64)
            reqs.append(dist.irecv(top, up))
        if down < world and min(bounds[down + 1], rows) > min(bounds[down], rows) and hi > lo:
            reqs.append(dist.isend(torch.from_numpy(local[-2].copy()), down))
            bot = torch.zeros(cols, dtype=torch.float64)
            reqs.append(dist.irecv(bot, down))
        for r in reqs:
            r.wait()
        if up >= 0 and hi > lo:
            local[0] = top.numpy()
        if down < world and hi > lo and min(bounds[down + 1], rows) > min(bounds[down], rows):
            local[-1] = bot.numpy()
        y_local = _apply(local)[1:-1]
        # the dot product of the slice, summed over ranks like dist_reduce does
        part = torch.tensor([float((x[lo:hi] * y_local).sum())], dtype=torch.float64)
        dist.all_reduce(part)
        y = _apply(x)
        assert np.allclose(y_local, y[lo:hi], rtol=0, atol=1e-12)
        assert abs(part.item() - float((x * y).sum())) < 1e-8 * abs(float((x * y).sum()))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_world_size_2_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(r, "ok") for r in range(world)], out
