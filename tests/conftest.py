"""Shared fixtures.  Tests marked `gpu` need a B200 (they call libsatfill.so through the C-ABI); everything else runs
on CPU: the oracle against the reference's golden vectors, the host-side logic and the ABI surface."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def port():
    import oracle

    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle

    return oracle.ref()  # None on machines without oracle/_ref (no /root/reference to build it from)


@pytest.fixture(scope="session")
def small_cases():
    return dict(np.load(os.path.join(GOLDEN, "small_cases.npz")))


@pytest.fixture(scope="session")
def c1_scene():
    d = dict(np.load(os.path.join(GOLDEN, "c1_scene.npz")))
    shape = tuple(int(v) for v in d["full_shape"])
    d["full_mask"] = np.unpackbits(d["full_mask_bits"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
    n = d["crop_b04"].shape[0]
    d["crop_mask"] = np.unpackbits(d["crop_mask_bits"])[: n * n].reshape(n, n).astype(bool)
    return d


@pytest.fixture(scope="session")
def c1_full(c1_scene):
    """BASELINE.json configs[0] / [1] at full size: the reference's sample scene with the reference's own converged Eigen
    solves (oracle/make_golden_full.py): unknown pixels only, 16.16 fixed point, delta-coded."""
    z = np.load(os.path.join(GOLDEN, "c1_full.npz"))
    scale = float(z["scale"])
    shape = tuple(int(v) for v in z["shape"])

    def unpack(planes):  # byte-transposed int32 second differences of 16.16 fixed point (oracle/make_golden_full.py)
        d2 = np.ascontiguousarray(planes.T).view(np.int32).reshape(-1).astype(np.int64)
        return np.cumsum(np.cumsum(d2)).astype(np.float64) / scale

    def unpack_band(planes):  # byte-transposed int16 row deltas of the uint16 band
        d = np.ascontiguousarray(planes.T).view(np.int16).reshape(shape).astype(np.int32)
        return np.cumsum(d, axis=1).astype(np.uint16).astype(np.float64)  # deltas wrap modulo 2^16

    full = c1_scene["full_mask"]
    lmask = full.copy()
    lmask[0, :] = lmask[-1, :] = False
    lmask[:, 0] = lmask[:, -1] = False
    return {"b04": unpack_band(z["b04_d"]), "b08": unpack_band(z["b08_d"]), "mask": full, "laplace_mask": lmask,
            "laplace_unknowns": unpack(z["laplace_unknowns_d"]), "laplace_iters": int(z["laplace_iters"]),
            "poisson_minus_guidance": [unpack(d) for d in z["poisson_minus_guidance_d"]], "poisson_iters": [int(i) for i in z["poisson_iters"]],
            "storage_abs_error": 0.5 / scale}  # fmt: skip


@pytest.fixture(scope="session")
def ctx():
    """One library context per test session (GPU tests only)."""
    import satellite_approximation_b200 as sab

    c = sab.Context(0)
    yield c
    c.close()


def rel_max_abs(x, x_ref, mask):
    """The parity measure of BASELINE.json: max |x - x_ref| / max |x_ref| over the unknowns."""
    a = np.asarray(x)[mask]
    b = np.asarray(x_ref)[mask]
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
