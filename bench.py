#!/usr/bin/env python
"""bench.py -- the Laplace / Poisson fill path on B200, measured on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c1|...]

A "step" is one pass of the hot path over one synthetic scene: mask -> unknown set and tile list, right-hand side,
preconditioned CG to the stop rule, filled pixels written in place.  Default workload (N = 1): configs[2] of
BASELINE.json, the configuration north_star quotes its target on -- a synthetic 10980 x 10980, 13-band Sentinel-2
tile whose mask is SURVEY.md 8d's "70th percentile of Gaussian-filtered (sigma = 40 px) white noise" (synth.cloud_mask:
window-reproducible, so the CPU reference arm solves a crop of the VERY SAME scene), Laplace fill to a 1e-6 relative
residual.

`value`   : device-resident (inputs in HBM when the timed region starts), CUDA events on the launching stream.  At N > 1
            every rank fills its own tile (scenes are independent: no data-path collective), `value` = all ranks'
            unknown pixels / max-over-ranks device time; `config.per_rank` lists every rank's time and CG iterations.
`e2e`     : the same metric through the host-pointer C-ABI call (sa_laplace_fill / sa_poisson_blend) on pinned host
            buffers; everything that crosses PCIe does so inside the timed region (`e2e.transfer` says how).
`e2e_dropin`: the call a user of the reference makes -- satellite_approximation.filling_missing_portions_smooth_boundaries /
            blend_images_poisson through the pybind11 `_core`, ordinary (pageable) numpy arrays, DEFAULTS ONLY (Laplace:
            epsilon tolerance and 2N iterations like the reference, laplace.cpp:113-114; src/main.cpp:49-58).
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration, accumulated inside the timed region
            (sa_options.profile), against MEASURED_PEAKS.json (`frac`) and against the 8 TB/s north_star names
            (`frac_of_nominal`); `step_frac` = algorithmic bytes of ALL solver kernels of the step / the whole step's time.
`cpu_baseline`: the reference's CPU path (oracle/_ref = its arithmetic on its vendored Eigen when that was built, else
            the plain-C port) on a crop of the same scene, in the three modes of BASELINE.md section 3.
At N > 1 the line also carries (north_star item 4):
`config.row_decomposed`: ONE system -- a configs[4]-shaped contiguous hole -- split by rows over the N ranks (csrc/dist.cu:
            halo rows and dot products over the library's own NCCL communicator); strong scaling (same hole at every N).
`config.one_tile_strong`: the 13 bands of rank 0's tile dealt round-robin to the N ranks (no collective).
`dist_parity`: the row-decomposed solve and the region sharding checked in-process against the single-GPU solve and the
            committed golden of the reference's Eigen (satellite_approximation_b200/distcheck.py); the run FAILS if not ok.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unknown_pixels_solved_per_s_at_1e-6_rel_residual"
UNIT = "px/s"
NOMINAL_HBM_GBS = 8000.0  # the figure north_star names; roofline.frac is against the measured copy bandwidth

WORKLOADS = {
    # name: rows, cols, bands, cover, sigma of the cloud field (px), problem
    "c3": dict(rows=10980, cols=10980, bands=13, cover=0.30, sigma=40.0, problem="laplace",
               desc="synthetic 10980x10980 13-band Sentinel-2 tile, 30% cloud-like mask, Laplace fill"),
    "c3-poisson": dict(rows=10980, cols=10980, bands=13, cover=0.30, sigma=40.0, problem="poisson",
                       desc="synthetic 10980x10980 13-band tile, 30% cloud-like mask, Poisson blend"),
    # round 1's mask (bicubically upsampled 48-pixel noise: blobs of about one tile), kept for comparison
    "c3-bicubic48": dict(rows=10980, cols=10980, bands=13, cover=0.30, cell=48, problem="laplace",
                         desc="synthetic 10980x10980 13-band tile, 30% mask of ~48 px blobs (round-1 generator), Laplace fill"),
    # the easy variant SURVEY.md 8d asks to report separately: iid Bernoulli(0.3) is sub-percolation (tiny components)
    "c3-iid": dict(rows=10980, cols=10980, bands=13, cover=0.30, problem="laplace", iid=True,
                   desc="synthetic 10980x10980 13-band tile, 30% iid Bernoulli mask (tiny components), Laplace fill"),
    "c1": dict(rows=1697, cols=1284, bands=5, cover=0.29, sigma=100.0, problem="laplace",
               desc="1697x1284 5-band scene (test_data/2019-05-22 shape), 29% mask, Laplace fill"),
    # ONE system shared by all ranks: split by rows, halo rows + dot products over NCCL (strong scaling)
    "c5": dict(rows=20000, cols=20000, bands=1, cover=1.0, problem="laplace", distributed=True,
               desc="single 20000x20000 contiguous hole, row-decomposed Laplace solve with halo exchange"),
    # 64 scenes x ~156 regions as one 8 x 8 mosaic: every region is its own linear system, the block-sparse tile list
    # batches them (only tiles that hold an unknown are ever visited)
    "c4": dict(rows=16384, cols=16384, bands=4, cover=None, problem="poisson", regions=10000,
               desc="batched many-small-holes: 10k independent cloud regions across 64 2048x2048 scenes (one mosaic), "
                    "4 bands, Poisson fill"),
    "small": dict(rows=2048, cols=2048, bands=4, cover=0.30, sigma=40.0, problem="laplace",
                  desc="2048x2048 4-band tile, 30% cloud-like mask, Laplace fill"),
}  # fmt: skip
CROP_ORIGIN = (4096, 4096)  # where the CPU arms cut their crop out of the scene


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SATFILL_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int)
    ap.add_argument("--cols", type=int)
    ap.add_argument("--bands", type=int)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--precond", default=os.environ.get("SATFILL_PRECOND", "multigrid"), choices=["jacobi", "multigrid"])
    ap.add_argument("--mg-variant", default=os.environ.get("SATFILL_MG_VARIANT", "rb32"), choices=["rb32", "jacobi64", "rb32_cta"])
    ap.add_argument("--cg-variant", type=int, default=int(os.environ.get("SATFILL_CG_VARIANT", "0")), choices=[0, 1])
    ap.add_argument("--check-every", type=int, default=0)
    ap.add_argument("--mask", default="", help="diagnostic mask patterns: full | tilecheck | halfrows | halfcols | tilecheck64")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the row-decomposed / one-tile / parity sub-records")
    ap.add_argument("--cpu-crop", type=int, default=768, help="edge of the crop the CPU baseline solves")
    ap.add_argument("--hole", type=int, default=20000, help="edge of the row-decomposed hole (configs[4]: 20000)")
    ap.add_argument("--rank-seeds", action="store_true",
                    help="N > 1: every rank builds a scene of its own (seed 2 + 17 rank) instead of its own copy of THE benchmark tile; "
                         "the masks then need 9 or 10 CG iterations and the slowest rank sets the step (config.per_rank shows it)")
    return ap.parse_args()


def workload(args):
    w = dict(WORKLOADS[args.workload])
    for k in ("rows", "cols", "bands"):
        if getattr(args, k):
            w[k] = getattr(args, k)
    return w


# ---- clocks -----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()  # fmt: skip
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}  # fmt: skip


def bind_to_gpu_numa_node(local: int, world: int) -> str:
    """Run this rank's host threads -- and so place its pinned buffers -- on the NUMA node its GPU hangs off, the way a
    launcher would with numactl; within the node the ranks take disjoint slices of the cores, so that the helper threads
    of eight ranks do not pile onto the same few."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = sorted(os.sched_getaffinity(0))
        where = "numa: single node"
        if node >= 0:
            on_node = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                on_node.update(range(int(a), int(b or a) + 1))
            if on_node & set(cpus):
                cpus = sorted(on_node & set(cpus))
                where = f"numa node {node}"
        per = max(len(cpus) // max(world, 1), 1)
        mine = cpus[(local * per) % len(cpus):][:per] or cpus
        os.sched_setaffinity(0, set(mine))
        return f"{where}, {len(mine)} cpus ({mine[0]}-{mine[-1]})"
    except Exception as e:  # noqa: BLE001 -- best effort: the bench runs unbound
        return f"numa: unbound ({type(e).__name__})"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- the scene (both arms) -------------------------------------------------------------------------------------------------
def scene_seed(rank: int) -> int:
    return 2 + 17 * rank


def crop_inputs(w, edge, nbands, rank=0):
    """numpy: the crop the CPU arms solve -- a window of rank `rank`'s scene (same generator, same seed, same pixels; the
    crop is its own image, so its border ring is cleared / known)."""
    from satellite_approximation_b200 import synth

    n = min(edge, w["rows"], w["cols"])
    r0 = min(CROP_ORIGIN[0], w["rows"] - n)
    c0 = min(CROP_ORIGIN[1], w["cols"] - n)
    if w.get("iid"):
        mask = synth.bernoulli_mask(n, n, cover=w["cover"], seed=2)
    elif w.get("regions"):  # the same density of regions as the workload
        mask = synth.region_mask(n, n, max(1, int(w["regions"] * n * n / (w["rows"] * w["cols"]))), seed=3)
    elif w.get("cell"):
        mask = synth.blob_mask(n, n, cover=w["cover"], sigma=w["cell"] / 3.0, seed=2)
    else:
        mask = synth.cloud_mask(n, n, cover=w["cover"], sigma=w["sigma"], seed=scene_seed(rank), row0=r0, col0=c0)
    bands = [synth.scene_band(n, n, seed=100 + b + 1000 * rank, row0=r0, col0=c0, total_rows=w["rows"], total_cols=w["cols"])
             for b in range(nbands)]  # fmt: skip
    return mask, bands, (r0, c0, n)


def device_inputs(w, rank, dev, nb=None, seed_rank=None):
    """torch, in HBM: rank `rank`'s scene."""
    import torch

    from satellite_approximation_b200 import synth

    rows, cols = w["rows"], w["cols"]
    nb = w["bands"] if nb is None else nb
    sr = rank if seed_rank is None else seed_rank
    if w.get("distributed"):
        # one hole covering everything but a one-pixel ring; every rank builds the same scene and owns a band of rows
        mask = torch.ones((rows, cols), dtype=torch.uint8, device=dev)
        sr = 0
    elif w.get("iid"):
        gen = torch.Generator(device=dev).manual_seed(scene_seed(sr))
        mask = (torch.rand((rows, cols), generator=gen, device=dev) < w["cover"]).to(torch.uint8)
    elif w.get("regions"):
        grid = 8
        mask = torch.from_numpy(synth.scene_mosaic_mask(rows // grid, grid, w["regions"], seed=3 + 7 * sr).view(np.uint8)).to(dev)
    elif w.get("cell"):
        mask = synth.torch_blob_mask(rows, cols, cover=w["cover"], cell=w["cell"], seed=scene_seed(sr), device=dev)
    else:
        mask = synth.torch_cloud_mask(rows, cols, cover=w["cover"], sigma=w["sigma"], seed=scene_seed(sr), device=dev)
    mask[0, :] = 0
    mask[-1, :] = 0
    mask[:, 0] = 0
    mask[:, -1] = 0
    return mask


def device_band(w, b, rank, dev):
    from satellite_approximation_b200 import synth

    return synth.torch_scene_band(w["rows"], w["cols"], seed=100 + b + 1000 * rank, device=dev)


# ---- CPU baseline (the checker, timed; never on the product path) --------------------------------------------------------
def cpu_solve(w, tol, edge, nbands, threads, steps=1, inputs=None):
    """The reference's CPU path on a crop of the scene, `nbands` bands on `threads` host threads (bands are independent:
    one band per thread -- ctypes releases the GIL; inside one band the reference's solve is single-threaded as shipped,
    SURVEY.md 2.1).  tol = None: the reference's own default (Laplace: epsilon, 2N iterations)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    from satellite_approximation_b200 import synth

    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    eng = ref if ref is not None else oracle.port()
    mask, bands, win = inputs if inputs is not None else crop_inputs(w, edge, nbands)
    bands = bands[:nbands]
    poisson = w["problem"] == "poisson"
    guides = [synth.second_date(b, seed=b_i) for b_i, b in enumerate(bands)] if poisson else None
    kw = {} if tol is None else {"tol": tol}

    def one(b):
        if poisson:
            _, st = eng.poisson_blend([bands[b]], [guides[b]], mask, **kw)
            return st[0]
        if kind == "reference":
            _, st = eng.laplace_fill(bands[b], mask, **kw)  # the reference's bounding-box system (laplace.cpp:31-120)
        else:
            _, st = eng.laplace_fill(bands[b], mask, mode=0, **kw)
        return st

    times, iters = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=max(1, min(threads, nbands))) as ex:
            sts = list(ex.map(one, range(nbands)))
        times.append(time.perf_counter() - t0)
        iters = max(s.iterations for s in sts)
    unknowns = int(mask.sum()) * nbands
    dt = statistics.median(times)
    n = win[2]
    return {
        "value": unknowns / dt, "unit": UNIT, "cores": max(1, min(threads, nbands)), "kind": kind,
        "sample": f"{n}x{n} crop at ({win[0]}, {win[1]}) of the workload's own scene (same generator and seed as rank 0), "
                  f"{nbands} band(s) one per host thread, {int(mask.sum())} unknowns/band, "
                  f"tol {'reference default' if tol is None else format(tol, 'g')}, {iters} CG iterations, {dt:.2f} s/step",
        "seconds_per_step": dt, "unknowns_per_step": unknowns, "cg_iterations": iters,
    }  # fmt: skip


def cpu_baseline_modes(w, tol, edge):
    """BASELINE.md section 3 / SURVEY.md 8d: (i) faithful -- one thread, the reference's defaults; (ii) tolerance-matched
    -- one thread at the bench's tolerance; (iii) best-effort N-core -- one band per host core at the bench's tolerance.
    The headline `value` is (iii); all three are reported."""
    threads = os.cpu_count() or 1
    nb = max(1, min(w["bands"], threads))
    inputs = crop_inputs(w, edge, nb)
    poisson = w["problem"] == "poisson"
    keep = ("value", "cores", "cg_iterations", "seconds_per_step")
    ncore = cpu_solve(w, tol, edge, nb, threads, inputs=inputs)
    matched = cpu_solve(w, tol, edge, 1, 1, inputs=inputs)
    faithful = matched if poisson and tol == 1e-6 else cpu_solve(w, None, edge, 1, 1, inputs=inputs)
    out = {k: ncore[k] for k in ("value", "unit", "cores", "kind", "sample")}
    out["mode"] = "ncore"
    out["modes"] = {"faithful": {k: faithful[k] for k in keep}, "tolerance_matched": {k: matched[k] for k in keep},
                    "ncore": {k: ncore[k] for k in keep}}  # fmt: skip
    out["host_cpus"] = threads
    return out


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nb = max(1, min(w["bands"], threads))
    # bounded so that warmup + steps end within a few minutes: time one step on the bench's own crop, shrink it if needed
    edge = min(args.cpu_crop, w["rows"], w["cols"])
    runs = max(args.steps, 1) + min(args.warmup, 1)
    probe = cpu_solve(w, args.tol, edge, nb, threads)
    if probe["seconds_per_step"] * runs > 240.0:
        edge = int(max(256, edge * (240.0 / (probe["seconds_per_step"] * runs)) ** 0.5 * 0.9))
        probe = cpu_solve(w, args.tol, edge, nb, threads)
    res = cpu_solve(w, args.tol, edge, nb, threads, steps=max(args.steps, 1)) if runs > 2 else probe
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "problem": w["problem"], "tolerance": args.tol, "mode": "ncore: one band per host core"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ---- the B200 arm ---------------------------------------------------------------------------------------------------------
KERNEL_NAMES_RB = ["cg_direction (k_direction2)", "cg_update (k_update2)", "mg coarse tail (k_rb_tail / k_rb_coarsest)",
                   "mg_transfer (unused on this path)", "mg_down level 0 (k_rb_down)", "mg_up level 0 (k_rb_up)",
                   "mg_down coarse levels", "mg_up coarse levels"]  # fmt: skip
KERNEL_NAMES_J64 = ["cg_direction (k_direction)", "cg_update (k_update)", "mg_sweep (k_mg_smooth)",
                    "mg_transfer (k_mg_residual/restrict/prolong)", "mg_down level 0 (k_mg_down)", "mg_up level 0 (k_mg_up)",
                    "mg_down coarse levels", "mg_up coarse levels"]  # fmt: skip
# algorithmic bytes per unknown of the level the launch runs on:
#   float red-black cycle: CG keeps its search direction in float and writes a float copy of r for the cycle --
#     direction R z 4 + R p 4 + W p 4 + mask 1 = 13;  update R p 4 + R/W x 16 + R/W r 16 + W rf 4 + mask 1 = 41;
#     down R b 4 + W b_c 4/4 = 5;  up R b 4 + R e_c 4/4 + W x 4 = 9  (the red half of the pre-smoothed iterate is
#     recomputed from b in the ascent, not carried through HBM; the first-generation kernels, --mg-variant rb32_cta, write
#     and read it: 7 and 11);  coarse levels also read the 1 / diagonal plane of the boundary-corrected operator: + 4
#   double Jacobi cycle: direction 25, single sweep 25, single transfer ~19, down 18, up 26
BYTES_RB = [13.0, 41.0, 22.0, 19.0, 5.0, 9.0, 9.0, 13.0]  # [2]: the coarse tail, descent 9 + ascent 13 per unknown of its levels
BYTES_RB_CTA = [13.0, 41.0, 9.0, 19.0, 7.0, 11.0, 11.0, 15.0]  # first-generation cycle kernels (the red half plane travels)
BYTES_J64 = [25.0, 41.0, 25.0, 19.0, 18.0, 26.0, 18.0, 26.0]
NK = 8


def rb_tables(rb):
    """Names and algorithmic bytes per unknown of the kernel classes of the red-black path.  The strip CG applies x += alpha p
    every other pass (k_update2, XM; SATFILL_DEFER_X=0 switches it off): the even passes leave x alone (41 - 16 = 25 B, class
    3), the odd ones add the steps of both (41 + R p_prev 4 = 45 B, class 1)."""
    names, bpu = list(KERNEL_NAMES_RB), list(BYTES_RB_CTA if rb == "cta" else BYTES_RB)
    if os.environ.get("SATFILL_DEFER_X", "1") != "0":
        names[1], bpu[1] = "cg_update, two steps into x (k_update2 XM=2)", 45.0
        names[3], bpu[3] = "cg_update, x left alone (k_update2 XM=1)", 25.0
    return names, bpu


class Timed:
    """K steps of a resident scene, timed with CUDA events on the launching stream; per-class kernel times accumulated
    from sa_options.profile."""

    def __init__(self, scene, mask, opts, stream):
        self.scene, self.mask, self.opts, self.stream = scene, mask, opts, stream
        self.kms, self.kn, self.ku = [0.0] * NK, [0] * NK, [0] * NK
        self.iters, self.st, self.ms, self.launches = [], None, 0.0, 0

    def step(self):
        self.scene.set_mask(self.mask)  # forces the whole path: indexing, hierarchy, right-hand side, solve, write-back
        return self.scene.solve(**self.opts)

    def run(self, steps, warmup, barrier):
        import torch

        for _ in range(warmup):
            self.st = self.step()
        barrier()
        l0 = self.scene.ctx.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            st = self.st = self.step()
            self.iters.append(max(s["iterations"] for s in st))
            for c in range(NK):
                self.kms[c] += st[0]["kernel_ms"][c]
                self.kn[c] += st[0]["kernel_launches"][c]
                self.ku[c] += st[0]["kernel_units"][c]
        e1.record(self.stream)
        self.launches = self.scene.ctx.kernel_launches - l0  # kernels of this library launched inside the timed region
        barrier()
        self.ms = e0.elapsed_time(e1)
        return self.ms


def kernel_table(t: Timed, rb, unit_scale=1.0):
    names, bpu = rb_tables(rb) if rb else (KERNEL_NAMES_J64, BYTES_J64)
    tab = {}
    for c in range(NK):
        if t.kn[c]:
            tab[names[c]] = {"ms": t.kms[c], "launches": t.kn[c], "bytes_per_unknown": bpu[c],
                             "GBps": (bpu[c] * t.ku[c] * unit_scale / (t.kms[c] * 1e-3) / 1e9) if t.kms[c] else None}  # fmt: skip
    return tab


def run_b200(args, w):
    import torch
    import torch.distributed as dist

    import satellite_approximation_b200 as sab
    from satellite_approximation_b200 import multi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local, world) if world > 1 else "numa: not bound (one rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows, cols, nb = w["rows"], w["cols"], w["bands"]
    poisson = w["problem"] == "poisson"
    problem = sab.POISSON if poisson else sab.LAPLACE
    precond = sab.MULTIGRID if args.precond == "multigrid" else sab.JACOBI
    # One explicit (non-default) stream for the synthetic data, the library and the timing events.  The legacy default
    # stream has handle 0, which sa_create reads as "no stream given": the library would then run on its own
    # non-blocking stream, unordered with torch's work.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = sab.Context(local, stream=stream.cuda_stream)
    one_system = bool(w.get("distributed"))
    if world > 1:
        ctx.dist_init_torch()  # the library's own NCCL communicator (row decomposition: csrc/dist.cu)

    # weak scaling: every rank fills its own copy of the benchmark tile (configs[2] names ONE tile: seed 2), so that the
    # per-N values differ by what running N GPUs at once costs and not by which masks the other ranks drew
    scene_rank = rank if args.rank_seeds else 0
    mask = device_inputs(w, scene_rank, dev)
    bands = [device_band(w, b, 0 if one_system else scene_rank, dev) for b in range(nb)]
    guides = [0.9 * bands[(b + 1) % nb] + 37.0 for b in range(nb)] if poisson else None
    if args.mask:  # diagnostic patterns: how the kernels' throughput depends on the shape of the unknown set
        rr = torch.arange(rows, device=dev)[:, None]
        cc = torch.arange(cols, device=dev)[None, :]
        pat = {"full": lambda: (rr >= 0) & (cc >= 0),
               "tilecheck": lambda: (((rr // 32) + (cc // 32)) % 2 == 0),
               "tilecheck64": lambda: (((rr // 32) + (cc // 64)) % 2 == 0),
               "halfrows": lambda: (cc % 32 < 16) & (rr >= 0),
               "halfcols": lambda: (rr % 32 < 16) & (cc >= 0)}[args.mask]()  # fmt: skip
        mask = pat.to(torch.uint8).contiguous()
        mask[0, :] = 0
        mask[-1, :] = 0
        mask[:, 0] = 0
        mask[:, -1] = 0
        del rr, cc, pat
    scene = ctx.scene(problem, rows, cols, nb)
    for b in range(nb):
        scene.set_band(b, bands[b])
        if poisson:
            scene.set_guidance(b, guides[b])
    if one_system and world > 1:
        scene.set_distributed(True)
    variant = {"rb32": sab.MG_RB32, "jacobi64": sab.MG_JACOBI64, "rb32_cta": sab.MG_RB32_CTA}[args.mg_variant]
    torch.cuda.synchronize()  # inputs resident in HBM before anything is timed
    opts = dict(tolerance=args.tol, precond=precond, profile=True, mg_variant=variant, cg_variant=args.cg_variant)
    if args.check_every:
        opts["check_every"] = args.check_every

    sampler = ClockSampler(local)
    sampler.start()
    timed = Timed(scene, mask, opts, stream)
    ms = timed.run(args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    launches = timed.launches
    st = timed.st
    # HBM held by this process while the resident scene exists (the library's planes + the bench's own copies of the inputs)
    free_b, total_b = torch.cuda.mem_get_info(dev)
    hbm_used_gb = (total_b - free_b) / 1e9
    unit_scale = 1.0  # the library's per-launch units are the unknowns THIS rank processed (row-decomposed solves included)
    unknowns = st[0]["unknowns"]
    ok = all(s["status"] == sab.SA_OK for s in st)
    worst_err = max(s["error"] for s in st)
    ms_max, total_units = multi.reduce_step(ms, float(unknowns * nb), one_system, dev)
    value = total_units * args.steps / (ms_max * 1e-3)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "ms_per_step": ms / max(args.steps, 1), "cg_iterations": timed.iters,
                "unknowns_per_band": unknowns, "solve_ms_last": st[0]["solve_ms"], "setup_ms_last": st[0]["setup_ms"]}  # fmt: skip
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---- roofline of the dominant kernel (by accumulated event time inside the timed region)
    rb = args.precond == "multigrid" and args.mg_variant != "jacobi64"
    if rb and args.mg_variant == "rb32_cta":
        rb = "cta"
    names, bpu = rb_tables(rb) if rb else (KERNEL_NAMES_J64, BYTES_J64)
    kms, kn, ku = timed.kms, timed.kn, [u * unit_scale for u in timed.ku]
    dom = max(range(NK), key=lambda c: kms[c])
    peak, peak_src = peaks()
    # DRAM traffic of the dominant kernel: bytes per unknown-band measured by one `ncu --set full` capture of the same
    # kernel on this workload (profiles/traffic.json, written by profiles/summarize.py traffic), scaled to this launch
    traffic_src = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(["k_direction2", "k_update2", None, "k_update2_x_left_alone" if rb else None, "k_rb_down", "k_rb_up", None, None][dom] or "")
        if ent and dom == 1 and rb and names[1] != KERNEL_NAMES_RB[1] and "float, 2>" not in ent.get("kernel", ""):
            ent = None  # a capture of the undeferred kernel does not describe this one
        if ent and args.workload == tj.get("workload", "c3") and not args.mask:
            traffic_src = ent
    except Exception:
        pass
    roof = None
    if kn[dom] > 0 and kms[dom] > 0:
        # per launch: algorithmic bytes = bytes/unknown x (unknown-bands the launches of this class processed / launches)
        units = ku[dom] / kn[dom]
        achieved = bpu[dom] * units / (kms[dom] / kn[dom] * 1e-3) / 1e9
        step_bytes = sum(bpu[c] * ku[c] for c in range(NK))
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal": achieved / NOMINAL_HBM_GBS, "nominal_peak": NOMINAL_HBM_GBS,
                "traffic": traffic_src["dram_bytes_per_unit"] * units if traffic_src else None,
                "traffic_source": (f"ncu dram__bytes_read+write of {traffic_src['kernel']}: {traffic_src['dram_bytes_per_unit']:.2f} B "
                                   f"per unknown-band ({traffic_src['capture']}), scaled to this launch's units") if traffic_src else None,
                "peak_source": peak_src,
                "avg_launch_ms": kms[dom] / kn[dom], "launches": kn[dom],
                "share_of_step": kms[dom] / ms if ms > 0 else None,
                "algorithmic_bytes_per_launch": bpu[dom] * units,
                "bytes_per_unknown": bpu[dom],
                # the whole step against the roofline: algorithmic bytes of every solver kernel launched in the timed region
                # (set-up, indexing and scrub kernels are in the time but not in the bytes) / the step's device time
                "step_GBps": step_bytes / (ms * 1e-3) / 1e9, "step_frac": step_bytes / (ms * 1e-3) / 1e9 / peak,
                "step_frac_of_nominal": step_bytes / (ms * 1e-3) / 1e9 / NOMINAL_HBM_GBS,
                "kernel_time_share_of_step": sum(kms) / ms if ms > 0 else None,
                "all_kernels": {k: dict(v, frac=(v["GBps"] / peak if v["GBps"] else None)) for k, v in kernel_table(timed, rb, unit_scale).items()}}  # fmt: skip

    # ---- end to end through the host-pointer C-ABI entry point (the resident scene is released first: the host-pointer
    # entry point keeps its own scene, and two 13-band scenes with solver work space do not fit one GPU together)
    scene.close()
    e2e = dropin = None
    h_arrays = None
    if not args.no_e2e and not one_system:
        e2e, h_arrays = run_e2e(args, w, ctx, sab, mask, bands, guides, world, dev, barrier, numa)
    bands.clear()
    if guides:
        guides.clear()
    torch.cuda.empty_cache()
    if not args.no_dropin and not one_system and world == 1 and h_arrays is not None:
        dropin = run_dropin(args, w, ctx, h_arrays)
    h_arrays = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_modes(w, args.tol, args.cpu_crop)

    # ---- N > 1: the multi-GPU shapes that are not replicas
    extra = {}
    rc = 0
    if world > 1 and not args.no_multi and not one_system:
        del mask
        torch.cuda.empty_cache()
        extra["one_tile_strong"] = run_one_tile_strong(args, w, ctx, sab, rank, world, dev, stream, barrier, multi)
        extra["row_decomposed"] = run_row_decomposed(args, ctx, sab, rank, world, dev, stream, barrier, multi, peak)
        from satellite_approximation_b200 import distcheck

        l0 = ctx.kernel_launches
        parity = distcheck.run_all(ctx, world, rank)
        parity["gpu_launches"] = ctx.kernel_launches - l0
        extra["dist_parity"] = parity
        if not parity["ok"]:
            rc = 3
    elif world == 1 and not args.no_multi and not one_system and not args.mask and args.workload == "c3":
        # the N = 1 point of the strong-scaling curve of the row-decomposed hole (same solve, one rank, no exchange)
        del mask
        torch.cuda.empty_cache()
        extra["row_decomposed"] = run_row_decomposed(args, ctx, sab, rank, world, dev, stream, barrier, multi, peak)

    if rank == 0:
        cfg = {"workload": w["desc"], "problem": w["problem"], "rows": rows, "cols": cols, "bands": nb,
               "mask": ("SURVEY 8d: Gaussian-filtered (sigma = %g px) white noise thresholded at the analytic %g quantile, border ring "
                        "cleared (synth.torch_cloud_mask, seed %s)" % (w["sigma"], 1 - w["cover"], "2 + 17 rank" if args.rank_seeds else "2 on every rank")) if w.get("sigma") and not args.mask else (args.mask or "see workload"),
               "setup_ms": st[0]["setup_ms"], "solve_ms": st[0]["solve_ms"], "hbm_used_gb": hbm_used_gb,
               "dtype_note": ("f64 CG iterate, residual, operator and dot products; f32 search direction and multigrid preconditioner"
                              if rb else "f64 throughout"),
               "unknowns_per_band": unknowns, "tolerance": args.tol, "precond": args.precond,
               "mg_variant": args.mg_variant if args.precond == "multigrid" else None,
               "cg_iterations": timed.iters, "cg_iterations_per_band": [s["iterations"] for s in st], "converged": ok,
               "worst_rel_residual": worst_err, "per_rank": per_rank,
               "l2": "inputs larger than L2 (no flush needed)" if rows * cols * 8 * nb > 2.6e8 else
                     "scene fits L2; mask re-upload + re-index between steps, no explicit flush",
               "parallelism": (f"one system split by rows over {world} GPU(s): NCCL halo rows + packed all-reduce of "
                               "the dot products, coarse multigrid levels replicated") if one_system else
                              f"{world} independent scene(s), one per GPU ({'own seeds' if args.rank_seeds else 'every rank its own copy of the benchmark tile'}), no collective"}  # fmt: skip
        for k in ("one_tile_strong", "row_decomposed"):
            if k in extra:
                cfg[k] = extra[k]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if one_system else "weak", "vs_baseline": None,
            "dtype": "f64",  # what the path computes in: CG iterate, residual, operator and dot products (config.dtype_note)
            "data": "synthetic", "config": cfg,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_dropin": dropin, "gpu_launches": launches, "clocks": clocks,
        }  # fmt: skip
        if "dist_parity" in extra:
            line["dist_parity"] = extra["dist_parity"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def run_e2e(args, w, ctx, sab, mask, bands, guides, world, dev, barrier, numa=""):
    import torch

    rows, cols, nb = w["rows"], w["cols"], w["bands"]
    poisson = w["problem"] == "poisson"
    h_mask = torch.empty((rows, cols), dtype=torch.uint8).pin_memory()
    h_mask.copy_(mask)
    h_bands = [torch.empty((rows, cols), dtype=torch.float64).pin_memory() for _ in range(nb)]
    for b in range(nb):
        h_bands[b].copy_(bands[b])
    h_guides = None
    if poisson:
        h_guides = [torch.empty((rows, cols), dtype=torch.float64).pin_memory() for _ in range(nb)]
        for b in range(nb):
            h_guides[b].copy_(guides[b])
    torch.cuda.synchronize()
    bands.clear()  # free the device copies: the e2e call starts from host memory
    if guides:
        guides.clear()
    torch.cuda.empty_cache()
    np_mask = h_mask.numpy()
    np_bands = [t.numpy() for t in h_bands]
    np_guides = [t.numpy() for t in h_guides] if poisson else None
    precond = sab.MULTIGRID if args.precond == "multigrid" else sab.JACOBI
    opts = dict(tolerance=args.tol, precond=precond, cg_variant=args.cg_variant,
                mg_variant={"rb32": sab.MG_RB32, "jacobi64": sab.MG_JACOBI64, "rb32_cta": sab.MG_RB32_CTA}[args.mg_variant])  # fmt: skip

    def call():
        # the filled pixels of the previous call are overwritten by the solver's own x0, so re-running on the same
        # buffers is the same work: known pixels are never modified
        if poisson:
            return ctx.poisson_blend(np_bands, np_guides, np_mask, **opts)
        return ctx.laplace_fill(np_bands, np_mask, **opts)

    n_e2e = max(1, args.steps)
    for _ in range(min(args.warmup, 2) or 1):  # warm-up: allocates the cached scene
        call()
    barrier()
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        st = call()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    my_dt = dt
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, my_dt / n_e2e)
    dt = float(t.item())
    unknowns = st[0]["unknowns"] * nb * world
    img_bytes = rows * cols * 8 * nb
    out = {"value": unknowns * n_e2e / dt, "unit": UNIT, "seconds_per_step": dt / n_e2e, "steps": n_e2e,
           "api": "sa_poisson_blend" if poisson else "sa_laplace_fill", "host_buffers": "pinned", "host_placement": numa,
           "seconds_per_step_per_rank": per_rank, "gpu_launches": ctx.kernel_launches - l0,
           "host_input_bytes": rows * cols + img_bytes * (2 if poisson else 1), "host_output_bytes": img_bytes}  # fmt: skip
    if ctx.last_fill_direct:
        # direct mode: kernels read / write the page-locked arrays in place.  What crosses PCIe per step: the mask, the
        # known pixels that border the unknown set (Poisson: plus g on the unknown set and around it), and the unknown
        # pixels on the way back -- counted here from the mask (the kernels move 16-byte pairs, so this is a lower bound)
        m = torch.from_numpy(np_mask).to(dev).bool()
        near = torch.zeros_like(m)
        near[1:, :] |= m[:-1, :]
        near[:-1, :] |= m[1:, :]
        near[:, 1:] |= m[:, :-1]
        near[:, :-1] |= m[:, 1:]
        ring = int((near & ~m).sum().item())
        unk = int(m.sum().item())
        del m, near
        out.update({"transfer": "direct (no image copies; set-up kernel reads, scatter kernel writes host memory)",
                    "h2d_bytes_per_step": rows * cols + 8 * nb * (ring + ((unk + ring) if poisson else 0)),
                    "d2h_bytes_per_step": 8 * nb * unk})  # fmt: skip
    else:
        out.update({"transfer": "copies (pipelined over three streams)", "h2d_bytes_per_step": out["host_input_bytes"],
                    "d2h_bytes_per_step": img_bytes})  # fmt: skip
    return out, (np_mask, np_bands, np_guides, (h_mask, h_bands, h_guides))


def run_dropin(args, w, ctx, h_arrays):
    """The reference's own Python surface (src/main.cpp:49-58) on ordinary numpy arrays, defaults only: one
    filling_missing_portions_smooth_boundaries call per band (the reference API takes one image; apply_laplace calls it
    once per channel, laplace.cpp:152-162), or one blend_images_poisson call for all bands.  Pageable memory: the library
    takes its copy path.  Laplace runs at the reference's default tolerance (epsilon), not at the bench's 1e-6."""
    import torch

    try:
        import satellite_approximation as sa
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"import satellite_approximation: {type(e).__name__}: {e}"}
    np_mask, np_bands, np_guides, _ = h_arrays
    rows, cols, nb = w["rows"], w["cols"], w["bands"]
    poisson = w["problem"] == "poisson"
    # what a user of the reference holds: a bool mask and float64 images in ordinary (pageable) memory, column-major like
    # the arrays the reference's own functions return (the pybind11 caster copies them into Eigen matrices:
    # src/main.cpp:49-54, both arguments .noconvert())
    mask = np.array(np_mask, dtype=bool, copy=True)
    unknowns = int(mask[1:-1, 1:-1].sum()) if not poisson else int(mask.sum())
    n_bands_timed = nb
    steps = max(1, min(args.steps, 2))

    def call(bands_np, guides_np):
        if poisson:
            return sa.blend_images_poisson(bands_np, guides_np, mask)
        return [sa.filling_missing_portions_smooth_boundaries(b, mask) for b in bands_np]

    bands_np = [np.asfortranarray(b) for b in np_bands]
    guides_np = [np.asfortranarray(g) for g in np_guides] if poisson else None
    mask = np.asfortranarray(mask)
    h_bytes = sum(b.nbytes for b in bands_np) * (2 if poisson else 1) + mask.nbytes * (1 if poisson else nb)
    call(bands_np[:1], guides_np[:1] if poisson else None)  # warm-up: library scene allocation, page faults of the casters
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    for _ in range(steps):
        out = call(bands_np, guides_np)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    changed = bool(np.any(out[0][mask] != bands_np[0][mask]))
    info = None
    try:
        from satellite_approximation import _core  # type: ignore

        p = _core.last_perf_info() if hasattr(_core, "last_perf_info") else None
        info = {"iterations": int(p["iterations"]), "error": float(p["error"]), "tolerance": float(p["tolerance"])} if p is not None else None
    except Exception:  # noqa: BLE001
        pass
    return {"value": unknowns * n_bands_timed / dt, "unit": UNIT, "seconds_per_step": dt, "steps": steps,
            "api": ("satellite_approximation.blend_images_poisson" if poisson else
                    f"satellite_approximation.filling_missing_portions_smooth_boundaries x {nb} bands") + f" ({getattr(sa, 'BACKEND', '?')} backend)",
            "host_buffers": "pageable numpy (F order, like the reference's own return values), copied by the pybind11 casters like the reference's",
            "options": "defaults only" + ("" if poisson else ": Laplace tolerance epsilon / 2N iterations (laplace.cpp:113-114), multigrid preconditioner"),
            "filled": changed, "last_perf_info": info, "h2d_bytes_per_step": h_bytes,
            "d2h_bytes_per_step": sum(b.nbytes for b in bands_np),
            "gpu_launches": int((ctx.kernel_launches - l0))}  # fmt: skip


def run_one_tile_strong(args, w, ctx, sab, rank, world, dev, stream, barrier, multi):
    """north_star's sentence as written: ONE 13-band tile on N GPUs.  The bands of rank 0's tile are dealt round-robin
    (bands share the mask and are independent right-hand sides: SURVEY.md 8e row 1; the mask is re-indexed by every rank);
    time = max over ranks of the device time of a step."""
    import torch

    nb = w["bands"]
    mine = multi.round_robin(nb, world, rank)
    rows, cols = w["rows"], w["cols"]
    mask = device_inputs(w, 0, dev)
    ms, iters, unknowns = 0.0, [], 0
    if mine:
        scene = ctx.scene(sab.LAPLACE if w["problem"] == "laplace" else sab.POISSON, rows, cols, len(mine))
        for i, b in enumerate(mine):
            band = device_band(w, b, 0, dev)
            scene.set_band(i, band)
            if w["problem"] == "poisson":
                scene.set_guidance(i, 0.9 * device_band(w, (b + 1) % nb, 0, dev) + 37.0)
            del band
        opts = dict(tolerance=args.tol, precond=sab.MULTIGRID if args.precond == "multigrid" else sab.JACOBI)
        t = Timed(scene, mask, opts, stream)
        steps = max(1, min(args.steps, 5))
        ms = t.run(steps, 2, barrier) / steps
        iters, unknowns = t.iters, t.st[0]["unknowns"]
        scene.close()
    else:
        barrier()
        barrier()
    del mask
    torch.cuda.empty_cache()
    ms_max, _ = multi.reduce_step(ms, 0.0, True, dev)
    u = torch.tensor([float(unknowns)], device=dev, dtype=torch.float64)
    import torch.distributed as dist

    dist.all_reduce(u, op=dist.ReduceOp.MAX)
    return {"what": f"the {nb} bands of ONE tile (rank 0's scene) dealt round-robin to {world} GPUs, no collective",
            "ms_per_tile": ms_max, "value": float(u.item()) * nb / (ms_max * 1e-3), "unit": UNIT, "bands_on_rank0": len(mine),
            "cg_iterations_rank0": iters}  # fmt: skip


def run_row_decomposed(args, ctx, sab, rank, world, dev, stream, barrier, multi, peak):
    """BASELINE.json configs[4]: a single contiguous hole as ONE linear system split by rows over the ranks (csrc/dist.cu).
    The same hole at every N: strong scaling.  value = unknowns / max-over-ranks device time of a step (index + hierarchy +
    plan + right-hand side + solve)."""
    import torch

    from satellite_approximation_b200 import synth

    n = args.hole
    w5 = dict(WORKLOADS["c5"], rows=n, cols=n)
    mask = device_inputs(w5, 0, dev)
    band = synth.torch_scene_band(n, n, seed=100, device=dev)
    scene = ctx.scene(sab.LAPLACE, n, n, 1)
    scene.set_band(0, band)
    del band
    if world > 1:
        scene.set_distributed(True)
    opts = dict(tolerance=args.tol, precond=sab.MULTIGRID if args.precond == "multigrid" else sab.JACOBI, profile=True)
    t = Timed(scene, mask, opts, stream)
    steps = max(1, min(args.steps, 5))
    ms = t.run(steps, 2, barrier) / steps
    launches = t.launches
    lo, hi, _ = scene.owned_rows() if world > 1 else (0, n, 0)
    st = t.st
    peer = ctx.dist_uses_peer_memory if world > 1 else False
    scene.close()
    del mask
    torch.cuda.empty_cache()
    ms_max, _ = multi.reduce_step(ms, 0.0, True, dev)
    tab = kernel_table(t, True)
    for v in tab.values():
        v["frac"] = v["GBps"] / peak if v["GBps"] else None
    return {"what": f"single {n}x{n} contiguous hole, one system split by rows over {world} GPU(s)"
                    + ((": halo rows and dot products over peer memory (CUDA IPC arenas, two kernels per exchange); NCCL for the plan "
                        "and the gather of the first replicated level" if peer else
                        ": halo rows and dot products over the library's NCCL communicator") if world > 1 else " (no exchange)"),
            "exchange": ("peer memory" if peer else "nccl") if world > 1 else None,
            "ms_per_solve": ms_max, "value": st[0]["unknowns"] / (ms_max * 1e-3), "unit": UNIT, "unknowns": st[0]["unknowns"],
            "cg_iterations": t.iters, "converged": all(s["status"] == sab.SA_OK for s in st),
            "worst_rel_residual": max(s["error"] for s in st), "setup_ms": st[0]["setup_ms"], "solve_ms": st[0]["solve_ms"],
            "rows_of_rank0": [int(lo), int(hi)], "gpu_launches_rank0": int(launches), "steps": steps,
            "kernels_rank0": tab}  # fmt: skip


def main():
    args = parse_args()
    w = workload(args)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
