"""Self-check of the multi-GPU paths, run in-process by `bench.py --gpus N` (so that the driver's scaling run exercises and
verifies csrc/dist.cu) and by tests/dist_worker.py (which adds the CPU oracle on top).

Two work shapes:
  * one linear system split by rows over the ranks of the library's NCCL communicator (`sa_dist_*`): every rank builds the
    same seeded scene, solves it as ONE system, gathers the bands, and compares them with the same solve done on its own
    GPU alone -- the same arithmetic up to the order of the partial sums of the dot products -- and, for the crop of the
    reference's sample scene, with the committed golden of the reference's own Eigen solve (tests/golden/c1_scene.npz;
    data only: nothing of oracle/ is imported here);
  * independent regions of one scene dealt out to the ranks (no collective in the solve, host-side merge).

Everything returns plain records; the callers decide what is fatal.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np

from . import LAPLACE, POISSON, MULTIGRID, JACOBI, SA_OK, Context, multi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_VS_SINGLE = 1e-8   # dist vs the same solve on one GPU (tight tolerances: only the summation order differs)
TOL_VS_GOLDEN = 1e-7   # dist vs the reference's converged Eigen solve


def rel(a: np.ndarray, b: np.ndarray, m: np.ndarray) -> float:
    return float(np.max(np.abs(a - b)[m]) / np.max(np.abs(b[m])))


def _solve(ctx: Context, problem, mask, bands, guides, distributed: bool, opts: dict):
    rows, cols = mask.shape
    nb = len(bands)
    sc = ctx.scene(problem, rows, cols, nb)
    try:
        sc.set_mask(mask)
        for b in range(nb):
            sc.set_band(b, bands[b])
            if problem == POISSON:
                sc.set_guidance(b, guides[b])
        owned = None
        if distributed:
            sc.set_distributed(True)
        st = sc.solve(**opts)
        if distributed:
            owned = sc.owned_rows()
            for b in range(nb):
                sc.allgather_band(b)
        return [sc.get_band(b) for b in range(nb)], st, owned
    finally:
        sc.close()


def golden_crop(golden_dir: Optional[str] = None):
    """The 320 x 320 crop of the reference's sample scene (test_data/2019-05-22, B04 + B08) with its mask, and the unknown
    pixels of the reference's converged Eigen Laplace solve of B04 (made by oracle/make_golden.py, committed)."""
    path = os.path.join(golden_dir or os.path.join(ROOT, "tests", "golden"), "c1_scene.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    n = z["crop_b04"].shape
    mask = np.unpackbits(z["crop_mask_bits"])[: n[0] * n[1]].reshape(n).astype(bool)
    return z["crop_b04"].astype(np.float64), z["crop_b08"].astype(np.float64), mask, z["laplace_unknowns"]


def row_decomposed_cases(ctx: Context, world: int, extra: Optional[Callable] = None, golden_dir: Optional[str] = None):
    """-> list of records {case, max_rel_vs_single, iterations_dist, iterations_single, ok, ...}.  `extra(name, problem,
    mask, bands, guides, dist_bands)` lets tests add their own comparison (the CPU oracle) and return a dict to merge."""
    cases = [
        ("laplace-mg", LAPLACE, (700, 900), dict(precond=MULTIGRID, tolerance=1e-11)),
        ("laplace-jacobi", LAPLACE, (300, 260), dict(precond=JACOBI, tolerance=1e-11)),
        ("poisson-mg", POISSON, (517, 389), dict(precond=MULTIGRID, tolerance=1e-11, max_iterations=10**5)),
        ("laplace-mg-hole", LAPLACE, (1500, 640), dict(precond=MULTIGRID, tolerance=1e-10)),
        ("laplace-mg-golden-crop", LAPLACE, None, dict(precond=MULTIGRID, tolerance=1e-11)),
    ]
    out = []
    for name, problem, shape, opts in cases:
        golden = None
        if shape is None:
            g = golden_crop(golden_dir)
            if g is None or g[2].shape[0] // 32 < world:
                continue
            bands, mask, golden = [g[0], g[1]], g[2], g[3]
            rows, cols = mask.shape
        else:
            rows, cols = shape
            if rows // 32 < world:
                continue
            bands = [synth.smooth_band(rows, cols, seed=11 + b) for b in range(2)]
            if name.endswith("hole"):  # one hole covering everything but a one-pixel ring (configs[4] in small)
                mask = np.zeros((rows, cols), bool)
                mask[1:-1, 1:-1] = True
            else:
                mask = synth.blob_mask(rows, cols, cover=0.4, sigma=9.0, seed=4, clear_border=problem == LAPLACE)
        guides = [synth.second_date(b, seed=3) for b in bands]
        single, st_s, _ = _solve(ctx, problem, mask, bands, guides, False, opts)
        dist_b, st_d, owned = _solve(ctx, problem, mask, bands, guides, True, opts)
        rec = {"case": name, "rows": rows, "cols": cols, "iterations_dist": [s["iterations"] for s in st_d],
               "iterations_single": [s["iterations"] for s in st_s], "owned_rows": list(owned[:2]) if owned else None}  # fmt: skip
        ok = all(s["status"] == SA_OK for s in st_d + st_s)
        worst = 0.0
        for b in range(len(bands)):
            worst = max(worst, rel(dist_b[b], single[b], mask))
            ok = ok and abs(st_d[b]["iterations"] - st_s[b]["iterations"]) <= 1
            ok = ok and bool(np.array_equal(dist_b[b][~mask], bands[b][~mask]))  # known pixels bit-identical
        rec["max_rel_vs_single"] = worst
        ok = ok and worst < TOL_VS_SINGLE
        if golden is not None:
            rec["max_rel_vs_reference_eigen_golden"] = float(np.max(np.abs(dist_b[0][mask] - golden)) / np.max(np.abs(golden)))
            ok = ok and rec["max_rel_vs_reference_eigen_golden"] < TOL_VS_GOLDEN
        if extra is not None:
            more = extra(name, problem, mask, bands, guides, dist_b) or {}
            ok = ok and bool(more.pop("ok", True))
            rec.update(more)
        rec["ok"] = bool(ok)
        out.append(rec)
    return out


def region_sharding_case(ctx: Context, world: int, rank: int):
    """Connected components of one scene packed onto the ranks (multi.pack_regions); every rank fills only its own
    components; the merged fills equal the whole-mask fill."""
    rows, cols, nb = 400, 520, 2
    mask = synth.region_mask(rows, cols, 40, area_lo=30.0, area_hi=4000.0, seed=9)
    bands = [synth.smooth_band(rows, cols, seed=21 + b) for b in range(nb)]
    lab, k = ctx.label_components(mask)
    shard, labels = multi.region_shard_mask(lab, k, world, rank)
    ok = bool(not (shard & ~mask).any())
    whole = [b.copy() for b in bands]
    ctx.laplace_fill(whole, mask, tolerance=1e-11, precond=MULTIGRID)
    part = [b.copy() for b in bands]
    if shard.any():
        ctx.laplace_fill(part, shard, tolerance=1e-11, precond=MULTIGRID)
    for b in range(nb):  # a rank touches only its own regions
        ok = ok and bool(np.array_equal(part[b][~shard], bands[b][~shard]))
    multi.merge_region_fills(part, shard)
    worst = max(rel(part[b], whole[b], mask) for b in range(nb))
    for b in range(nb):
        ok = ok and bool(np.array_equal(part[b][~mask], bands[b][~mask]))
    return {"case": "regions", "components": int(k), "mine": len(labels), "max_rel_vs_whole_mask": worst,
            "ok": bool(ok and worst < TOL_VS_SINGLE)}  # fmt: skip


def run_all(ctx: Context, world: int, rank: int, extra: Optional[Callable] = None) -> dict:
    """All checks on this rank, and-ed over the ranks (torch.distributed must be initialised, ctx.dist_init_torch() done).
    -> {"ok", "max_rel", "cases": [...]}"""
    import torch
    import torch.distributed as dist

    recs = row_decomposed_cases(ctx, world, extra)
    recs.append(region_sharding_case(ctx, world, rank))
    ok = all(r["ok"] for r in recs)
    worst = max([r.get("max_rel_vs_single", 0.0) for r in recs] + [r.get("max_rel_vs_whole_mask", 0.0) for r in recs]
                + [r.get("max_rel_vs_reference_eigen_golden", 0.0) for r in recs])  # fmt: skip
    dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([0.0 if ok else 1.0, worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"ok": bool(t[0].item() == 0.0), "max_rel": float(t[1].item()), "cases": recs,
            "tolerances": {"vs_single_gpu": TOL_VS_SINGLE, "vs_reference_eigen_golden": TOL_VS_GOLDEN}}  # fmt: skip
