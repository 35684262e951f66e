"""Worker of tests/test_host_api.py::test_cpp_shim_behaviour_on_the_cpu: started with LD_LIBRARY_PATH pointing at a
libsatfill.so built from tests/fake_satfill.c (the C-ABI answered by the oracle), it drives the pybind11 module
satellite_approximation._core -- i.e. the C++ `approx` shim -- and checks its host-side behaviour against the oracle."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    import oracle
    from satellite_approximation import _core as core
    from satellite_approximation_b200 import synth

    port = oracle.port()
    img = synth.smooth_band(50, 64, seed=1)
    mask = synth.blob_mask(50, 64, cover=0.3, sigma=4.0, seed=2)
    # Laplace: noconvert dtype rules, a new F-ordered array, argument untouched, known pixels bit-identical
    keep = img.copy()
    got = core.filling_missing_portions_smooth_boundaries(img, mask)
    want = port.laplace_fill(img, mask, mode=1)[0]
    assert got.flags.f_contiguous and np.array_equal(img, keep)
    assert np.array_equal(got[~mask], img[~mask]) and np.allclose(got, want, rtol=0, atol=1e-9 * np.abs(want).max())
    for bad in ((img.astype(np.float32), mask), (img, mask.astype(np.uint8))):
        try:
            core.filling_missing_portions_smooth_boundaries(*bad)
            raise SystemExit("noconvert was not enforced")
        except TypeError:
            pass
    try:
        core.filling_missing_portions_smooth_boundaries(img, mask[:, :-1])
        raise SystemExit("size mismatch did not throw")
    except RuntimeError:  # laplace.cpp:124-127
        pass
    same = core.filling_missing_portions_smooth_boundaries(img, np.zeros_like(mask))  # empty mask: laplace.cpp:41-44
    assert np.array_equal(same, img)
    # strided (non-contiguous) arguments go through the Eigen casters
    big = np.zeros((100, 128))
    big[::2, ::2] = img
    assert np.array_equal(core.filling_missing_portions_smooth_boundaries(big[::2, ::2], mask), got)
    # Poisson: list in, list out, defaults, failure leaves the inputs
    f = [img, synth.smooth_band(50, 64, seed=5)]
    g = [synth.second_date(x, seed=7 + i) for i, x in enumerate(f)]
    out = core.blend_images_poisson(f, g, mask)
    wantp = port.poisson_blend(f, g, mask, tol=1e-6)[0]
    assert len(out) == 2 and all(np.array_equal(o, w) for o, w in zip(out, wantp))
    out = core.blend_images_poisson(f, g, mask, tolerance=1e-13, max_iterations=2)  # poisson.cpp:263-269
    assert all(np.array_equal(o, x) for o, x in zip(out, f))
    out = core.blend_images_poisson(f, [x[:, :-1] for x in g], mask)  # poisson.cpp:154-157: log and return
    assert all(np.array_equal(o, x) for o, x in zip(out, f))
    # set_log_level is wired (src/main.cpp:30-34), the record of the last fill is readable, apply_laplace runs on numpy images
    assert core.get_log_level() == core.LogLevel.Warn
    core.set_log_level(core.LogLevel.Critical)
    assert core.get_log_level() == core.LogLevel.Critical
    core.set_log_level(core.LogLevel.Warn)
    core.blend_images_poisson(f, g, mask)
    info = core.last_perf_info()
    assert info["region_size"] == int(mask.sum()) and info["iterations"] >= 1 and info["tolerance"] == 1e-6, info
    rng = np.random.default_rng(11)
    bgr = rng.integers(0, 256, (40, 36, 3), dtype=np.uint8)
    marked = np.zeros_like(bgr)
    marked[5:20, 8:30, 2] = 255  # red, green stays 0: (R >= 220) & (G <= 150)
    al = core.apply_laplace(bgr, marked, 220.0)
    wal, wmask = oracle.apply_laplace(bgr, marked, 220.0, tol=1e-13)
    assert al.shape == (40, 36, 3) and wmask.sum() == 15 * 22 and np.max(np.abs(al - wal)) < 1e-6
    # connected components: the reference's own case (tests/approximation.h:55-75)
    m = np.zeros((10, 10), bool)
    m[1:3, 1:3] = True
    m[5:9, 5:7] = True
    lab, k = core.find_connected_components(m)
    assert k == 2 and (lab == 1).sum() == 4 and (lab == 2).sum() == 8 and lab.dtype == np.int32
    # offset / white-key overload against the dense restatement
    ins = [synth.smooth_band(40, 48, seed=b, lo=0.0, hi=1.0) for b in range(3)]
    rep = [synth.smooth_band(15, 17, seed=10 + b, lo=0.0, hi=0.9) for b in range(3)]
    key = np.zeros((15, 17), bool)
    key[0, :] = key[-1, :] = key[:, 0] = key[:, -1] = True
    key[5:8, 6:9] = True
    for ch in rep:
        ch[key] = 1.2
    got = core.blend_images_poisson_offset(ins, rep, 9, 11)
    want = oracle.poisson_offset_dense(ins, rep, 9, 11)
    region = np.zeros((40, 48), bool)
    region[9 : 9 + 15, 11 : 11 + 17] = ~key
    for b in range(3):
        assert np.max(np.abs(got[b] - want[b])[region]) < 1e-8 and np.array_equal(got[b][~region], ins[b][~region])
    same = core.blend_images_poisson_offset(ins, rep, 30, 7)  # out of bounds: returned unchanged (poisson.cpp:25-39)
    assert all(np.array_equal(s, a) for s, a in zip(same, ins))
    print("core-on-fake ok")


if __name__ == "__main__":
    main()
