"""Worker of tests/test_gpu_dist.py: one rank of a row-decomposed solve (launched by torch.distributed.run, one process
per GPU).  Every rank builds the same seeded scene, solves it as ONE system split by rows over NCCL, gathers the bands
and compares them with the same solve done on its own GPU alone and with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import satellite_approximation_b200 as sab  # noqa: E402
from satellite_approximation_b200 import synth  # noqa: E402


def rel(a, b, m):
    return float(np.max(np.abs(a - b)[m]) / np.max(np.abs(b[m])))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sab.Context(local)
    ctx.dist_init_torch()
    port = oracle.port()
    cases = [
        ("laplace-mg", sab.LAPLACE, (700, 900), dict(precond=sab.MULTIGRID, tolerance=1e-11)),
        ("laplace-jacobi", sab.LAPLACE, (300, 260), dict(precond=sab.JACOBI, tolerance=1e-11)),
        ("poisson-mg", sab.POISSON, (517, 389), dict(precond=sab.MULTIGRID, tolerance=1e-11, max_iterations=10**5)),
        ("laplace-mg-hole", sab.LAPLACE, (1500, 640), dict(precond=sab.MULTIGRID, tolerance=1e-10)),
    ]
    for name, problem, (rows, cols), opts in cases:
        nb = 2
        bands = [synth.smooth_band(rows, cols, seed=11 + b) for b in range(nb)]
        if name.endswith("hole"):  # one hole covering everything but a one-pixel ring (BASELINE.json configs[4] in small)
            mask = np.zeros((rows, cols), bool)
            mask[1:-1, 1:-1] = True
        else:
            mask = synth.blob_mask(rows, cols, cover=0.4, sigma=9.0, seed=4, clear_border=problem == sab.LAPLACE)
        guides = [synth.second_date(b, seed=3) for b in bands]
        outs = {}
        for mode in ("single", "dist"):
            sc = ctx.scene(problem, rows, cols, nb)
            sc.set_mask(mask)
            for b in range(nb):
                sc.set_band(b, bands[b])
                if problem == sab.POISSON:
                    sc.set_guidance(b, guides[b])
            if mode == "dist":
                sc.set_distributed(True)
            st = sc.solve(**opts)
            assert all(s["status"] == sab.SA_OK for s in st), (name, mode, [s["status"] for s in st])
            if mode == "dist":
                lo, hi, axis = sc.owned_rows()
                assert axis == 0 and 0 <= lo < hi <= rows
                for b in range(nb):
                    sc.allgather_band(b)
            outs[mode] = ([sc.get_band(b) for b in range(nb)], [s["iterations"] for s in st])
            sc.close()
        for b in range(nb):
            # the same arithmetic up to the order of the partial sums of the dot products
            assert rel(outs["dist"][0][b], outs["single"][0][b], mask) < 1e-8, (name, b)
            assert abs(outs["dist"][1][b] - outs["single"][1][b]) <= 1, (name, outs["dist"][1], outs["single"][1])
            assert np.array_equal(outs["dist"][0][b][~mask], bands[b][~mask])
        if rows * cols < 400000:
            if problem == sab.POISSON:
                want, _ = port.poisson_blend(bands, guides, mask, tol=1e-13, max_it=10**6)
            else:
                want = [port.laplace_fill(b_, mask, mode=1, tol=1e-13)[0] for b_ in bands]
            for b in range(nb):
                assert rel(outs["dist"][0][b], want[b], mask) < 1e-7, (name, b)
        if rank == 0:
            print(f"ok {name}: iterations dist {outs['dist'][1]} single {outs['single'][1]}", flush=True)
    # ---- independent regions of one scene dealt out to the ranks (no collective in the solve; host-side merge)
    from satellite_approximation_b200 import multi

    rows, cols, nb = 400, 520, 2
    mask = synth.region_mask(rows, cols, 40, area_lo=30.0, area_hi=4000.0, seed=9)
    bands = [synth.smooth_band(rows, cols, seed=21 + b) for b in range(nb)]
    lab, k = ctx.label_components(mask)
    shard, labels = multi.region_shard_mask(lab, k, world, rank)
    assert shard.any() and not (shard & ~mask).any() and len(labels) >= k // world - 1
    whole = [b.copy() for b in bands]
    ctx.laplace_fill(whole, mask, tolerance=1e-11, precond=sab.MULTIGRID)
    part = [b.copy() for b in bands]
    ctx.laplace_fill(part, shard, tolerance=1e-11, precond=sab.MULTIGRID)
    for b in range(nb):  # a rank touches only its own regions
        assert np.array_equal(part[b][~shard], bands[b][~shard])
    multi.merge_region_fills(part, shard)
    for b in range(nb):
        assert rel(part[b], whole[b], mask) < 1e-8 and np.array_equal(part[b][~mask], bands[b][~mask])
    if rank == 0:
        print(f"ok regions: {k} components over {world} ranks", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
