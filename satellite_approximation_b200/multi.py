"""Host-side logic of the multi-GPU paths, one process per GPU, torch.distributed for the plumbing.

Two work shapes:
  * independent systems (bands of a scene, scenes, connected regions): dealt out to ranks, NO data-path collective --
    `round_robin`, `pack_regions`;
  * one system split by rows (sa_dist_*, csrc/dist.cu): the library owns the NCCL communicator; the host language only
    moves the 128-byte id (`broadcast_bytes`) and may ask for the row partition (`row_partition`).
`reduce_step` is the timing rule of bench.py: device time = max over ranks, work = sum over ranks (or counted once when
all ranks share one system).  Everything here runs on CPU with the gloo backend as well (tests/test_multi_gloo.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

from . import _capi


def round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Items (bands, scenes) of rank `rank`: i with i % world == rank."""
    return list(range(rank, n_items, world))


def pack_regions(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Size-sorted greedy bin packing of independent regions (connected components are independent linear systems:
    4-connectivity, approx/utils.h:38-44) over `world` GPUs: largest first onto the least loaded rank.  Deterministic
    (ties broken by index), so every rank computes the same assignment without talking to the others."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        k = min(range(world), key=lambda r: (load[r], r))
        bins[k].append(i)
        load[k] += int(sizes[i])
    return bins


def row_partition(rows: int, world: int, levels: int | None = None) -> Tuple[List[int], int]:
    """(row boundaries (world + 1), multigrid levels split by rows) of the row decomposition (sa_dist_partition)."""
    lib = _capi.load()
    if levels is None:
        levels = int(lib.sa_dist_levels(int(rows), int(world)))
    out = (C.c_int64 * (world + 1))()
    if lib.sa_dist_partition(int(rows), int(world), int(levels), out) != 0:
        raise ValueError("row_partition: bad arguments")
    return list(out), levels


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """`payload` of rank `src` on every rank (the ncclUniqueId of sa_dist_unique_id).  Works on the nccl backend (a
    device tensor) and on gloo (a CPU tensor)."""
    import torch
    import torch.distributed as dist

    if device is None:
        device = torch.device("cpu")
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def reduce_step(ms: float, units: float, one_system: bool, device=None) -> Tuple[float, float]:
    """(max over ranks of the device time, units of work of the whole job): independent scenes add up, one shared
    system is counted once."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms), float(units)
    if device is None:
        device = torch.device("cpu")
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    u = torch.tensor([units], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if not one_system:
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def region_shard_mask(labels, num_labels: int, world: int, rank: int):
    """Independent regions of ONE scene over several GPUs (second row): connected components are independent
    linear systems, so rank `rank` fills the components that `pack_regions` deals to it and nothing else.  `labels` is
    the label image of sa_label_components (0 = valid pixel).  Returns (this rank's invalid mask, its component labels).
    The union of the ranks' fills is the fill of the whole mask; merging is a host-side gather of disjoint pixel sets
    (`merge_region_fills`)."""
    import numpy as np

    lab = np.asarray(labels)
    sizes = np.bincount(lab.ravel(), minlength=num_labels + 1)[1:]
    mine = pack_regions(sizes, world)[rank]
    keep = np.zeros(num_labels + 1, bool)
    keep[[i + 1 for i in mine]] = True
    return keep[lab], [i + 1 for i in mine]


def merge_region_fills(filled, shard_mask):
    """All ranks' fills of their own regions merged into every rank's `filled` arrays (a list of 2-D arrays, modified in
    place): each rank contributes the pixels of its shard mask; the sets are disjoint, so the merge is an all-gather of
    (mask, values) pairs and a scatter.  No collective touches the solve itself."""
    import numpy as np
    import torch.distributed as dist

    world = dist.get_world_size()
    mine = (np.asarray(shard_mask), [np.asarray(a)[shard_mask] for a in filled])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    for m, vals in parts:
        for a, v in zip(filled, vals):
            a[m] = v
    return filled
