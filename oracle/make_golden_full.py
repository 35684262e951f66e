"""TEST INFRASTRUCTURE ONLY -- tests/golden/c1_full.npz: BASELINE.json configs[0] / configs[1] at FULL size, from the
reference's own arithmetic.

Run here (the container that has /root/reference):   python oracle/make_golden_full.py        (about 3 minutes)

The reference's sample scene test_data/2019-05-22 (1697 x 1284; executables/laplace-main.cpp:34-40,
poisson-main.cpp:53-70), solved by oracle/_ref/libref_eigen.so = the reference's assembly (laplace.cpp:31-120,
poisson.cpp:145-290) executed by the reference's vendored Eigen:
  * Laplace, band B04, mask = (R >= 220) & (G <= 150) of selected_pixels.png with the image-border ring cleared
    (633 332 unknowns; SURVEY.md F5: the reference does not converge on the uncleared mask), Eigen defaults (epsilon);
  * Poisson, bands B04 / B08, guidance = synth.second_date of the other band, the uncleared mask (633 573 unknowns),
    tolerance 1e-10.
Stored: the two uint16 bands, and per solve the UNKNOWN pixels only, as 16.16 fixed point (absolute error 7.6e-6 on values
up to 10^4, i.e. < 1e-8 of the value range -- two orders below the tightest parity bar) second-differenced along the raster order and
byte-transposed so that deflate gets them down to about 1 MB each; the bands are row-delta-coded the same way.  The Poisson solves are stored as x - g (g = the seeded guidance image the test
regenerates): the blend is g plus a harmonic correction, which is smooth where x itself carries g's noise.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from satellite_approximation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SCENE = "/root/reference/test_data/2019-05-22"
SCALE = 65536.0


def pack(values: np.ndarray) -> np.ndarray:
    """16.16 fixed point, second differences along the raster order of the unknowns (the fills are smooth: a harmonic
    function's second difference is small), bytes transposed so that deflate sees the quiet high bytes together."""
    q = np.rint(values * SCALE).astype(np.int64)
    d2 = np.diff(np.diff(q, prepend=0), prepend=0)
    assert np.abs(d2).max() < 2**31
    return np.ascontiguousarray(d2.astype(np.int32).view(np.uint8).reshape(-1, 4).T)


def unpack(planes: np.ndarray) -> np.ndarray:
    d2 = np.ascontiguousarray(planes.T).view(np.int32).reshape(-1).astype(np.int64)
    return np.cumsum(np.cumsum(d2)).astype(np.float64) / SCALE


def pack_band(band: np.ndarray) -> np.ndarray:
    d = np.diff(band.astype(np.int32), axis=1, prepend=0).astype(np.int16)
    return np.ascontiguousarray(d.view(np.uint8).reshape(-1, 2).T)


def unpack_band(planes: np.ndarray, shape) -> np.ndarray:
    d = np.ascontiguousarray(planes.T).view(np.int16).reshape(shape).astype(np.int32)
    return np.cumsum(d, axis=1).astype(np.uint16)


def main() -> None:
    import cv2

    oracle.build(ref=True)
    ref = oracle.ref()
    assert ref is not None, "oracle/_ref/libref_eigen.so could not be built (no /root/reference?)"
    png = cv2.imread(os.path.join(SCENE, "selected_pixels.png"), cv2.IMREAD_UNCHANGED)
    full_mask = (png[..., 2] >= 220) & (png[..., 1] <= 150)  # laplace.cpp:141-146, red threshold 220 (laplace-main.cpp:38)
    assert full_mask.sum() == 633573, full_mask.sum()
    b04 = cv2.imread(os.path.join(SCENE, "B04.tif"), cv2.IMREAD_UNCHANGED)
    b08 = cv2.imread(os.path.join(SCENE, "B08.tif"), cv2.IMREAD_UNCHANGED)
    assert b04.dtype == np.uint16 and b04.shape == full_mask.shape == (1697, 1284)
    lmask = full_mask.copy()
    lmask[0, :] = lmask[-1, :] = False
    lmask[:, 0] = lmask[:, -1] = False
    assert lmask.sum() == 633332, lmask.sum()
    t0 = time.time()
    lap, st = ref.laplace_fill(b04.astype(np.float64), lmask, tol=0.0, max_it=0)  # Eigen defaults: epsilon, 2N
    assert st.status == 0, st
    print(f"laplace: {st.iterations} iterations, error {st.error:.3e}, {time.time() - t0:.1f} s", flush=True)
    f = [b04.astype(np.float64), b08.astype(np.float64)]
    g = [synth.second_date(f[1], seed=0), synth.second_date(f[0], seed=1)]
    t0 = time.time()
    poi, pst = ref.poisson_blend(f, g, full_mask, tol=1e-10, max_it=10**6)
    assert all(s.status == 0 for s in pst), pst
    print(f"poisson: {[s.iterations for s in pst]} iterations, {time.time() - t0:.1f} s", flush=True)
    lap_u = np.ascontiguousarray(lap)[lmask]
    poi_u = [np.ascontiguousarray(o)[full_mask] - g[b][full_mask] for b, o in enumerate(poi)]
    for v in [lap_u] + poi_u:
        assert np.max(np.abs(unpack(pack(v)) - v)) < 1e-5
    assert np.array_equal(unpack_band(pack_band(b04), b04.shape), b04)
    np.savez_compressed(
        os.path.join(OUT, "c1_full.npz"),
        b04_d=pack_band(b04), b08_d=pack_band(b08), shape=np.array(b04.shape, np.int64), scale=np.float64(SCALE),
        laplace_unknowns_d=pack(lap_u), laplace_iters=np.int64(st.iterations),
        poisson_minus_guidance_d=np.stack([pack(v) for v in poi_u]), poisson_iters=np.array([s.iterations for s in pst], np.int64),
        poisson_tol=np.float64(1e-10),
    )
    print("c1_full.npz", os.path.getsize(os.path.join(OUT, "c1_full.npz")), "bytes")


if __name__ == "__main__":
    main()
