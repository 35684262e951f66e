#!/usr/bin/env python
"""bench.py -- the Laplace / Poisson fill path on B200, measured on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c1|...]

A "step" is one pass of the hot path over one synthetic scene: mask -> unknown set and tile list, right-hand side,
preconditioned CG to the stop rule, filled pixels written in place.  Default workload (N = 1): configs[2] of
BASELINE.json, the configuration north_star quotes its target on -- a synthetic 10980 x 10980, 13-band Sentinel-2
tile with a 30 % cloud-like mask, Laplace fill to a 1e-6 relative residual.  At N > 1 every rank fills its own tile
(scenes are independent: no data-path collective), `value` = all ranks' unknown pixels / max-over-ranks device time.

`value`   : device-resident (inputs in HBM when the timed region starts), CUDA events on the launching stream.
`e2e`     : the same metric through the host-pointer C-ABI call (sa_laplace_fill / sa_poisson_blend) on pinned host
            buffers; everything that crosses PCIe does so inside the timed region.  Pinned buffers put the library in
            its direct mode (kernels read the known pixels bordering the unknown set from host memory and store the
            unknown pixels back: no image copies); `e2e.transfer` says which way the call went and `h2d / d2h
            _bytes_per_step` count what really moved.
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration, accumulated inside the timed region
            (sa_options.profile), against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (oracle/_ref = the reference's arithmetic on its vendored Eigen when that was built, else
            the plain-C port) on a bounded crop of the same workload, on this box's host cores.
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

WORKLOADS = {
    # name: rows, cols, bands, cover, blob cell (px), problem
    "c3": dict(rows=10980, cols=10980, bands=13, cover=0.30, cell=48, problem="laplace",
               desc="synthetic 10980x10980 13-band Sentinel-2 tile, 30% cloud-like mask, Laplace fill"),
    "c3-poisson": dict(rows=10980, cols=10980, bands=13, cover=0.30, cell=48, problem="poisson",
                       desc="synthetic 10980x10980 13-band tile, 30% cloud-like mask, Poisson blend"),
    # the easy variant SURVEY.md 8d asks to report separately: iid Bernoulli(0.3) is sub-percolation (tiny components)
    "c3-iid": dict(rows=10980, cols=10980, bands=13, cover=0.30, cell=0, problem="laplace", iid=True,
                   desc="synthetic 10980x10980 13-band tile, 30% iid Bernoulli mask (tiny components), Laplace fill"),
    "c1": dict(rows=1697, cols=1284, bands=5, cover=0.29, cell=160, problem="laplace",
               desc="1697x1284 5-band scene (test_data/2019-05-22 shape), 29% mask, Laplace fill"),
    # ONE system shared by all ranks: split by rows, halo rows + dot products over NCCL (strong scaling)
    "c5": dict(rows=20000, cols=20000, bands=1, cover=1.0, cell=0, problem="laplace", distributed=True,
               desc="single 20000x20000 contiguous hole, row-decomposed Laplace solve with halo exchange"),
    # 64 scenes x ~156 regions as one 8 x 8 mosaic: every region is its own linear system, the block-sparse tile list
    # batches them (only tiles that hold an unknown are ever visited)
    "c4": dict(rows=16384, cols=16384, bands=4, cover=None, cell=0, problem="poisson", regions=10000,
               desc="batched many-small-holes: 10k independent cloud regions across 64 2048x2048 scenes (one mosaic), "
                    "4 bands, Poisson fill"),
    "small": dict(rows=2048, cols=2048, bands=4, cover=0.30, cell=48, problem="laplace",
                  desc="2048x2048 4-band tile, 30% cloud-like mask, Laplace fill"),
}  # fmt: skip


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
    ap.add_argument("--mg-variant", default=os.environ.get("SATFILL_MG_VARIANT", "rb32"), choices=["rb32", "jacobi64"])
    ap.add_argument("--cg-variant", type=int, default=int(os.environ.get("SATFILL_CG_VARIANT", "0")), choices=[0, 1])
    ap.add_argument("--check-every", type=int, default=0)
    ap.add_argument("--mask", default="", help="diagnostic mask patterns: full | tilecheck | halfrows | halfcols | tilecheck64")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-crop", type=int, default=768, help="edge of the crop the CPU baseline solves")
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


def bind_to_gpu_numa_node(local: int) -> str:
    """Run this rank's host threads -- and so place its pinned buffers -- on the NUMA node its GPU hangs off, the way a
    launcher would with numactl: with eight ranks moving 25 GB each per step, host memory on the wrong socket makes
    every PCIe transfer cross the inter-socket link."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as e:  # noqa: BLE001 -- best effort: the bench runs unbound
        return f"numa: unbound ({type(e).__name__})"
    return "numa: unbound"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU baseline (the checker, timed; never on the product path) --------------------------------------------------------
def cpu_baseline(w, tol, crop, threads, steps=1):
    """Reference CPU path on a bounded crop of the same workload.  Bands are independent, so they are solved one per
    host thread (ctypes releases the GIL); inside one band the reference's solve is single-threaded as shipped."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    from satellite_approximation_b200 import synth

    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    eng = ref if ref is not None else oracle.port()
    n = min(crop, w["rows"], w["cols"])
    if w.get("iid"):
        mask = synth.bernoulli_mask(n, n, cover=w["cover"], seed=2)
    elif w.get("regions"):  # the same density of regions as the workload
        mask = synth.region_mask(n, n, max(1, int(w["regions"] * n * n / (w["rows"] * w["cols"]))), seed=3)
    else:
        mask = synth.blob_mask(n, n, cover=w["cover"], sigma=w["cell"] / 3.0, seed=2)
    nb = max(1, min(w["bands"], threads))
    bands = [synth.smooth_band(n, n, seed=100 + b) for b in range(nb)]
    poisson = w["problem"] == "poisson"
    guides = [synth.second_date(b, seed=b_i) for b_i, b in enumerate(bands)] if poisson else None

    def one(b):
        if poisson:
            if kind == "reference":
                _, st = eng.poisson_blend([bands[b]], [guides[b]], mask, tol=tol)
            else:
                _, st = eng.poisson_blend([bands[b]], [guides[b]], mask, tol=tol)
            return st[0]
        if kind == "reference":
            _, st = eng.laplace_fill(bands[b], mask, tol=tol)  # the reference's bounding-box system (laplace.cpp:31-120)
        else:
            _, st = eng.laplace_fill(bands[b], mask, mode=0, tol=tol)
        return st

    times, iters = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=nb) as ex:
            sts = list(ex.map(one, range(nb)))
        times.append(time.perf_counter() - t0)
        iters = max(s.iterations for s in sts)
    unknowns = int(mask.sum()) * nb
    dt = statistics.median(times)
    return {
        "value": unknowns / dt, "unit": UNIT, "cores": nb, "kind": kind,
        "sample": f"{n}x{n} crop of the workload, {nb} band(s) one per host thread, {int(mask.sum())} unknowns/band, "
                  f"tol {tol:g}, {iters} CG iterations, {dt:.2f} s/step",
        "seconds_per_step": dt, "unknowns_per_step": unknowns,
    }  # fmt: skip


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded so that warmup + steps end within a few minutes: calibrate on a small crop, then pick the crop edge
    probe = cpu_baseline(w, args.tol, 256, threads)
    per_px = probe["seconds_per_step"] / max(probe["unknowns_per_step"] / probe["cores"], 1)
    budget = 150.0 / max(args.steps + min(args.warmup, 1), 1)
    edge = int(min(max((budget / max(per_px, 1e-12) / w["cover"]) ** 0.5 * 0.5, 256), args.cpu_crop, w["rows"]))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(w, args.tol, edge, threads)
    res = cpu_baseline(w, args.tol, edge, threads, steps=max(args.steps, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "problem": w["problem"], "tolerance": args.tol},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ---- the B200 arm ---------------------------------------------------------------------------------------------------------
def run_b200(args, w):
    import torch
    import torch.distributed as dist

    import satellite_approximation_b200 as sab
    from satellite_approximation_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "numa: not bound (one rank)"
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
    if one_system and world > 1:
        ctx.dist_init_torch()

    if one_system:
        # one hole covering everything but a one-pixel ring; every rank builds the same scene and owns a band of rows
        mask = torch.ones((rows, cols), dtype=torch.uint8, device=dev)
        mask[0, :] = 0
        mask[-1, :] = 0
        mask[:, 0] = 0
        mask[:, -1] = 0
        bands = [synth.torch_band(rows, cols, seed=100 + b, device=dev) for b in range(nb)]
    elif w.get("iid"):
        gen = torch.Generator(device=dev).manual_seed(2 + 17 * rank)
        mask = (torch.rand((rows, cols), generator=gen, device=dev) < w["cover"]).to(torch.uint8)
        mask[0, :] = 0
        mask[-1, :] = 0
        mask[:, 0] = 0
        mask[:, -1] = 0
        bands = [synth.torch_band(rows, cols, seed=100 + b + 1000 * rank, device=dev) for b in range(nb)]
    elif w.get("regions"):
        grid = 8
        mask = torch.from_numpy(synth.scene_mosaic_mask(rows // grid, grid, w["regions"], seed=3 + 7 * rank).view(np.uint8)).to(dev)
        bands = [synth.torch_band(rows, cols, seed=100 + b + 1000 * rank, device=dev) for b in range(nb)]
    else:
        # synthetic scene, built in HBM; every rank gets its own seed (independent scenes)
        mask = synth.torch_blob_mask(rows, cols, cover=w["cover"], cell=w["cell"], seed=2 + 17 * rank, device=dev)
        bands = [synth.torch_band(rows, cols, seed=100 + b + 1000 * rank, device=dev) for b in range(nb)]
    guides = [0.9 * bands[(b + 1) % nb] + 37.0 for b in range(nb)] if poisson else None
    if args.mask:  # diagnostic patterns: how the kernels' throughput depends on the shape of the unknown set
        rr = torch.arange(rows, device=dev)[:, None]
        cc = torch.arange(cols, device=dev)[None, :]
        pat = {"full": lambda: (rr >= 0) & (cc >= 0),
               "tilecheck": lambda: (((rr // 32) + (cc // 32)) % 2 == 0),
               "tilecheck64": lambda: (((rr // 32) + (cc // 64)) % 2 == 0),
               "halfrows": lambda: (cc % 32 < 16) & (rr >= 0),
               "halfcols": lambda: (rr % 32 < 16) & (cc >= 0)}[args.mask]()
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
    variant = sab.MG_RB32 if args.mg_variant == "rb32" else sab.MG_JACOBI64
    torch.cuda.synchronize()  # inputs resident in HBM before anything is timed
    opts = dict(tolerance=args.tol, precond=precond, profile=True, mg_variant=variant, cg_variant=args.cg_variant)
    if args.check_every:
        opts["check_every"] = args.check_every

    def step():
        scene.set_mask(mask)  # forces the whole path: indexing, (hierarchy,) right-hand side, solve, write-back
        return scene.solve(**opts)

    for _ in range(args.warmup):
        st = step()
    unknowns = st[0]["unknowns"] if args.warmup else None
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    NK = 8
    kms = [0.0] * NK
    kn = [0] * NK
    ku = [0] * NK
    iters = []
    for _ in range(args.steps):
        st = step()
        iters.append(max(s["iterations"] for s in st))
        for c in range(NK):
            kms[c] += st[0]["kernel_ms"][c]
            kn[c] += st[0]["kernel_launches"][c]
            ku[c] += st[0]["kernel_units"][c]
    e1.record(stream)
    barrier()
    if one_system and world > 1:
        # the library counts a launch's units as the whole system's unknowns: a rank processes its share of the rows
        lo, hi, _ = scene.owned_rows()
        ku = [u * (hi - lo) / rows for u in ku]
    clocks = sampler.stop()
    launches = ctx.kernel_launches - launches0
    ms = e0.elapsed_time(e1)
    unknowns = st[0]["unknowns"]
    ok = all(s["status"] == sab.SA_OK for s in st)
    worst_err = max(s["error"] for s in st)
    from satellite_approximation_b200 import multi

    ms_max, total_units = multi.reduce_step(ms, float(unknowns * nb), one_system, dev)
    value = total_units * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (by accumulated event time inside the timed region)
    rb = args.precond == "multigrid" and args.mg_variant == "rb32"
    names = ["cg_direction (k_direction)", "cg_update (k_update)", "mg_sweep (k_mg_smooth / k_rb_coarsest)",
             "mg_transfer (k_mg_residual/restrict/prolong)",
             "mg_down level 0 (k_rb_down)" if rb else "mg_down level 0 (k_mg_down)",
             "mg_up level 0 (k_rb_up)" if rb else "mg_up level 0 (k_mg_up)",
             "mg_down coarse levels", "mg_up coarse levels"]
    # algorithmic bytes per unknown of the level the launch runs on, fp64 CG vectors:
    #   direction: R z 8 (4: the float cycle's z) + R p 8 + W p 8 + R mask 1;  update: R p, x, r 24 + W x, r 16 + mask 1
    #   single smoother sweep: R x 8 + R b 8 + W x 8 + R mask 1 = 25;  single transfers ~19
    #   double Jacobi cycle: down R b 8 + W x 8 + W b_c 8/4 = 18;  up R x 8 + R b 8 + R e_c 8/4 + W x 8 = 26
    #   float red-black cycle: CG keeps its search direction in float and writes a float copy of r for the cycle --
    #     direction R z 4 + R p 4 + W p 4 + mask 1 = 13;  update R p 4 + R/W x 16 + R/W r 16 + W rf 4 + mask 1 = 41:
    #     down R b 4 + W x_red 4/2 + W b_c 4/4 = 7;  up R x_red 4/2 + R b 4 + R e_c 4/4 + W x 4 = 11;
    #     coarse levels also read the 1 / diagonal plane of the boundary-corrected operator: 11 and 15
    if rb:
        bytes_per_unknown = [13.0, 41.0, 9.0, 19.0, 7.0, 11.0, 11.0, 15.0]
    else:
        bytes_per_unknown = [25.0, 41.0, 25.0, 19.0, 18.0, 26.0, 18.0, 26.0]
    dom = max(range(NK), key=lambda c: kms[c])
    peak, peak_src = peaks()
    # DRAM traffic of the dominant kernel: bytes per unknown-band measured by one `ncu --set full` capture of the same
    # kernel on this workload (profiles/traffic.json, written by profiles/summarize.py traffic), scaled to this launch
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(["k_direction2", "k_update2", None, None, "k_rb_down", "k_rb_up", None, None][dom] or "")
        if ent and args.workload == tj.get("workload", "c3") and not args.mask:
            traffic_src = ent
    except Exception:
        pass
    roof = None
    if kn[dom] > 0 and kms[dom] > 0:
        # per launch: algorithmic bytes = bytes/unknown x (unknown-bands the launches of this class processed / launches)
        units = ku[dom] / kn[dom]
        achieved = bytes_per_unknown[dom] * units / (kms[dom] / kn[dom] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic_src["dram_bytes_per_unit"] * units if traffic_src else None,
                "traffic_source": (f"ncu dram__bytes_read+write of {traffic_src['kernel']}: {traffic_src['dram_bytes_per_unit']:.2f} B "
                                   f"per unknown-band ({traffic_src['capture']}), scaled to this launch's units") if traffic_src else None,
                "peak_source": peak_src,
                "avg_launch_ms": kms[dom] / kn[dom], "launches": kn[dom],
                "share_of_step": kms[dom] / ms if ms > 0 else None,
                "algorithmic_bytes_per_launch": bytes_per_unknown[dom] * units,
                "bytes_per_unknown": bytes_per_unknown[dom],
                "all_kernels": {names[c]: {"ms": kms[c], "launches": kn[c],
                                           "GBps": (bytes_per_unknown[c] * ku[c] / (kms[c] * 1e-3) / 1e9) if kms[c] else None}
                                for c in range(NK)}}  # fmt: skip

    # ---- end to end through the host-pointer C-ABI entry point (the resident scene is released first: the host-pointer
    # entry point keeps its own scene, and two 13-band scenes with solver work space do not fit one GPU together)
    scene.close()
    e2e = None
    if not args.no_e2e and not one_system:
        e2e = run_e2e(args, w, ctx, sab, mask, bands, guides, world, dev, barrier, numa)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(w, args.tol, args.cpu_crop, os.cpu_count() or 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if one_system else "weak", "vs_baseline": None,
            "dtype": "f64" + (" (CG iterate, residual, operator and dot products; float inside the multigrid preconditioner)" if rb else ""),
            "data": "synthetic",
            "config": {"workload": w["desc"], "problem": w["problem"], "rows": rows, "cols": cols, "bands": nb,
                       "setup_ms": st[0]["setup_ms"], "solve_ms": st[0]["solve_ms"],
                       "unknowns_per_band": unknowns, "tolerance": args.tol, "precond": args.precond,
                       "mg_variant": args.mg_variant if args.precond == "multigrid" else None,
                       "cg_iterations": iters, "cg_iterations_per_band": [s["iterations"] for s in st], "converged": ok, "worst_rel_residual": worst_err,
                       "l2": "inputs larger than L2 (no flush needed)" if rows * cols * 8 * nb > 2.6e8 else
                             "scene fits L2; mask re-upload + re-index between steps, no explicit flush",
                       "parallelism": (f"one system split by rows over {world} GPU(s): NCCL halo rows + packed all-reduce of "
                                       "the dot products, coarse multigrid levels replicated") if one_system else
                                      f"{world} independent scene(s), one per GPU, no collective"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    del mask
    torch.cuda.empty_cache()
    np_mask = h_mask.numpy()
    np_bands = [t.numpy() for t in h_bands]
    np_guides = [t.numpy() for t in h_guides] if poisson else None
    precond = sab.MULTIGRID if args.precond == "multigrid" else sab.JACOBI
    opts = dict(tolerance=args.tol, precond=precond, cg_variant=args.cg_variant,
                mg_variant=sab.MG_RB32 if args.mg_variant == "rb32" else sab.MG_JACOBI64)

    def call():
        # the filled pixels of the previous call are overwritten by the solver's own x0, so re-running on the same
        # buffers is the same work: known pixels are never modified
        if poisson:
            return ctx.poisson_blend(np_bands, np_guides, np_mask, **opts)
        return ctx.laplace_fill(np_bands, np_mask, **opts)

    n_e2e = max(1, min(args.steps, 3))
    call()  # warm-up: allocates the cached scene
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        st = call()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    unknowns = st[0]["unknowns"] * nb * world
    img_bytes = rows * cols * 8 * nb
    out = {"value": unknowns * n_e2e / dt, "unit": UNIT, "seconds_per_step": dt / n_e2e, "steps": n_e2e,
           "api": "sa_poisson_blend" if poisson else "sa_laplace_fill", "host_buffers": "pinned", "host_placement": numa,
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
    return out


def main():
    args = parse_args()
    w = workload(args)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
