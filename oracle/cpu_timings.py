"""The three CPU timings BASELINE.md section 3 asks for, on configs[0] / configs[1] of BASELINE.json (the reference's own sample
scene test_data/2019-05-22) -- TEST INFRASTRUCTURE: this runs the reference's arithmetic (oracle/_ref: the assembly of
laplace.cpp:31-120 / poisson.cpp:145-290 on the reference's vendored Eigen), never the product.

  1. faithful           1 thread, reference defaults: Laplace tol = epsilon, maxIter = 2N, x0 = 0;
                        Poisson tol = 1e-6, maxIter = n / 2, guess = replacement
  2. tolerance-matched  Laplace at 1e-6 (Poisson already is), 1 thread
  3. best-effort N-core Eigen's OpenMP SpMV on all cores (what `Eigen::setNbThreads(hardware_concurrency)` of
                        poisson-main.cpp:35-37 would give if OpenMP were linked into the library), and one band per core

    python oracle/cpu_timings.py [--quick] > profiles/<round>_cpu_timings.json

Needs /root/reference (the sample scene and the Eigen headers oracle/_ref is built from): it runs in the build container,
not on the GPU box; `bench.py`'s cpu_baseline / --impl reference legs are the in-run CPU numbers there.  `--quick` uses
the committed 320 x 320 crop (tests/golden/c1_scene.npz) instead of the full scene."""
from __future__ import annotations

import json
import os
import platform
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SCENE = "/root/reference/test_data/2019-05-22"


def load(quick: bool):
    from satellite_approximation_b200 import geotiff, synth  # the repo's own TIFF reader (no GDAL here)

    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_scene.npz")))
    if quick or not os.path.isdir(SCENE):
        n = d["crop_b04"].shape[0]
        mask = np.unpackbits(d["crop_mask_bits"])[: n * n].reshape(n, n).astype(bool)
        bands = [d["crop_b04"].astype(np.float64), d["crop_b08"].astype(np.float64)]
        what = f"{n}x{n} crop of test_data/2019-05-22 (tests/golden/c1_scene.npz)"
    else:
        shape = tuple(int(v) for v in d["full_shape"])
        mask = np.unpackbits(d["full_mask_bits"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
        mask[0, :] = mask[-1, :] = False  # border ring cleared: the reference has no valid answer on it
        mask[:, 0] = mask[:, -1] = False
        bands = [geotiff.TiffFile(os.path.join(SCENE, f"{b}.tif")).read_band(1).astype(np.float64)
                 for b in ("B02", "B03", "B04", "B08", "B11")]  # fmt: skip
        what = f"test_data/2019-05-22 {shape[0]}x{shape[1]}, selected_pixels.png mask with the border ring cleared"
    guides = [synth.second_date(bands[(i + 1) % len(bands)], seed=i) for i in range(len(bands))]
    return mask, bands, guides, what


def timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def main() -> None:
    import oracle

    quick = "--quick" in sys.argv
    ref = oracle.ref()
    if ref is None:
        raise SystemExit("oracle/_ref is not built (needs the reference's vendored Eigen)")
    mask, bands, guides, what = load(quick)
    n_unknown = int(mask.sum())
    cores = os.cpu_count() or 1
    nb = len(bands)
    rows = []

    def row(config, mode, threads, seconds, unknown_bands, iters, note=""):
        rows.append({"config": config, "mode": mode, "threads": threads, "seconds": round(seconds, 3),
                     "unknown_px_per_s": round(unknown_bands / seconds, 1), "cg_iterations": iters, "note": note})  # fmt: skip
        print(f"# {config:8s} {mode:26s} {threads:2d} thr  {seconds:8.2f} s  {unknown_bands / seconds:12.0f} px/s  {iters}",
              file=sys.stderr)  # fmt: skip

    # ---- configs[0]: Laplace fill of one band (B04) ----------------------------------------------------------------------
    b04 = bands[2] if nb >= 3 else bands[0]
    ref.set_threads(1)
    dt, (_, st) = timed(lambda: ref.laplace_fill(b04, mask, tol=0.0, max_it=0))
    row("laplace", "faithful (tol eps, 2N it)", 1, dt, n_unknown, st.iterations, f"estimated error {st.error:.2e}")
    dt, (_, st) = timed(lambda: ref.laplace_fill(b04, mask, tol=1e-6, max_it=0))
    row("laplace", "tolerance-matched (1e-6)", 1, dt, n_unknown, st.iterations)
    ref.set_threads(cores)
    dt, (_, st) = timed(lambda: ref.laplace_fill(b04, mask, tol=1e-6, max_it=0))
    row("laplace", "N-core (OpenMP SpMV)", cores, dt, n_unknown, st.iterations, "one band: only Eigen's SpMV is threaded")
    ref.set_threads(1)
    k = min(nb, cores)
    with ThreadPoolExecutor(k) as ex:
        dt, sts = timed(lambda: list(ex.map(lambda b: ref.laplace_fill(b, mask, tol=1e-6, max_it=0)[1], bands[:k])))
    row("laplace", "N-core (one band per core)", k, dt, n_unknown * k, max(s.iterations for s in sts), f"{k} bands")

    # ---- configs[1]: Poisson blend against a synthetic second date ----------------------------------------------------------
    dt, (_, st) = timed(lambda: ref.poisson_blend(bands[:1], guides[:1], mask, tol=1e-6))
    row("poisson", "faithful = matched (1e-6)", 1, dt, n_unknown, st[0].iterations, "one band")
    ref.set_threads(cores)
    dt, (_, st) = timed(lambda: ref.poisson_blend(bands[:1], guides[:1], mask, tol=1e-6))
    row("poisson", "N-core (OpenMP SpMV)", cores, dt, n_unknown, st[0].iterations, "one band")
    ref.set_threads(1)
    with ThreadPoolExecutor(k) as ex:
        dt, sts = timed(lambda: list(ex.map(
            lambda i: ref.poisson_blend([bands[i]], [guides[i]], mask, tol=1e-6)[1][0], range(k))))  # fmt: skip
    row("poisson", "N-core (one band per core)", k, dt, n_unknown * k, max(s.iterations for s in sts), f"{k} bands")

    print(json.dumps({
        "what": "reference CPU path (oracle/_ref: reference assembly on the reference's vendored Eigen 3.4.90, g++ -O2)",
        "input": what, "unknowns_per_band": n_unknown, "host": {"cpu": platform.processor() or platform.machine(),
                                                               "model": _cpu_model(), "logical_cores": cores},
        "timings": rows}, indent=1))  # fmt: skip


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


if __name__ == "__main__":
    main()
