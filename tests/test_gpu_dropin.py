"""GPU: the drop-in surfaces above the C-ABI -- the C++ `approx` shim behind the pybind11 module
`satellite_approximation._core` (cpp/, what a user of the reference's Python package imports) and the offset / white-key
Poisson overload (poisson.cpp:21-143) in both host languages -- against the oracle."""
from __future__ import annotations

import numpy as np
import pytest
from conftest import rel_max_abs

import satellite_approximation_b200 as sab
from satellite_approximation_b200 import synth

pytestmark = pytest.mark.gpu


def _core():
    try:
        from satellite_approximation import _core
    except ImportError:
        pytest.skip("satellite_approximation._core is not built (make -C cpp pybind needs Eigen headers)")
    return _core


def _paste_case(seed=0, rows=60, cols=72, R=23, C=31):
    rng = np.random.default_rng(seed)
    ins = [synth.smooth_band(rows, cols, seed=seed + b, lo=0.0, hi=1.0) for b in range(3)]
    rep = [synth.smooth_band(R, C, seed=seed + 10 + b, lo=0.0, hi=0.9) for b in range(3)]
    key = synth.blob_mask(R, C, cover=0.45, sigma=3.0, seed=seed + 5, clear_border=False)
    key[0, :] = key[-1, :] = True  # the chair on its white background: the key surrounds the object ...
    key[:, 0] = True
    key[R // 2, C - 1] = False     # ... except where the object touches the edge of the replacement image
    for ch in rep:
        ch[key] = 1.0 + 0.5 * rng.random(int(key.sum()))  # truncates to 1 in every channel: the white key
    return ins, rep, key


def test_offset_overload_python_mirror_vs_dense_oracle(ctx):
    import oracle

    ins, rep, key = _paste_case()
    assert np.array_equal(sab.valid_pixel_mask(rep), ~key)
    for r0, c0 in ((5, 7), (0, 0), (60 - 23, 72 - 31)):
        want = oracle.poisson_offset_dense(ins, rep, r0, c0)
        got = [a.copy() for a in ins]
        sab.blend_images_poisson_offset(got, rep, r0, c0)
        region = np.zeros(ins[0].shape, bool)
        region[r0 : r0 + 23, c0 : c0 + 31] = ~key
        for b in range(3):
            assert rel_max_abs(got[b], want[b], region) < 1e-8
            assert np.array_equal(got[b][~region], ins[b][~region])  # only the unknowns are written (poisson.cpp:126-139)
    # the three bounds checks log and return with the input untouched (poisson.cpp:25-39)
    for bad in ((-1, 0), (0, 72), (50, 7), (5, 60)):
        got = [a.copy() for a in ins]
        sab.blend_images_poisson_offset(got, rep, *bad)
        assert all(np.array_equal(g, a) for g, a in zip(got, ins))
    big = [np.ones((80, 90)) for _ in range(3)]
    got = [a.copy() for a in ins]
    sab.blend_images_poisson_offset(got, big, 0, 0)
    assert all(np.array_equal(g, a) for g, a in zip(got, ins))


def test_offset_overload_cpp_shim_vs_dense_oracle():
    import oracle

    core = _core()
    ins, rep, key = _paste_case(seed=3)
    want = oracle.poisson_offset_dense(ins, rep, 9, 11)
    got = core.blend_images_poisson_offset(ins, rep, 9, 11)
    region = np.zeros(ins[0].shape, bool)
    region[9 : 9 + 23, 11 : 11 + 31] = ~key
    for b in range(3):
        assert rel_max_abs(got[b], want[b], region) < 1e-8
        assert np.array_equal(got[b][~region], ins[b][~region])
    same = core.blend_images_poisson_offset(ins, rep, 50, 7)  # out of bounds: returned unchanged
    assert all(np.array_equal(g, a) for g, a in zip(same, ins))


def test_pybind_module_is_a_drop_in(port):
    """satellite_approximation (the reference's package name) over the C++ shim: same names, noconvert rules, return
    layout and results as the ctypes mirror and the oracle."""
    import satellite_approximation as sa

    core = _core()
    assert sa.BACKEND == "pybind11"
    rows, cols = 70, 90
    img = np.asfortranarray(synth.smooth_band(rows, cols, seed=1))
    mask = np.asfortranarray(synth.blob_mask(rows, cols, cover=0.35, sigma=4.0, seed=2))
    want, _ = port.laplace_fill(img, mask, mode=1, tol=1e-13)
    core.set_laplace_options(tolerance=1e-12, max_iterations=0, multigrid=True)
    got = sa.filling_missing_portions_smooth_boundaries(img, mask)
    assert got.flags["F_CONTIGUOUS"] and got.dtype == np.float64 and got is not img
    assert rel_max_abs(got, want, mask) < 1e-8 and np.array_equal(got[~mask], img[~mask])
    with pytest.raises(TypeError):
        sa.filling_missing_portions_smooth_boundaries(img.astype(np.float32), mask)  # noconvert (src/main.cpp:49-54)
    with pytest.raises(RuntimeError):
        sa.filling_missing_portions_smooth_boundaries(img, np.asfortranarray(mask[:-1]))  # laplace.cpp:124-127
    # the reference's own preconditioner as the opt-in, then back to the defaults: a plain drop-in call (no knobs:
    # epsilon tolerance, 2N iterations, laplace.cpp:113-114) runs the multigrid path and lands on the same fill
    core.set_laplace_options(tolerance=0.0, max_iterations=0, multigrid=False)
    gotj = sa.filling_missing_portions_smooth_boundaries(img, mask)
    core.set_laplace_options(tolerance=0.0, max_iterations=0, multigrid=True)
    gotd = sa.filling_missing_portions_smooth_boundaries(img, mask)
    assert rel_max_abs(gotj, want, mask) < 1e-9 and rel_max_abs(gotd, want, mask) < 1e-9
    f = [synth.smooth_band(rows, cols, seed=5 + b) for b in range(2)]
    g = [synth.second_date(b, seed=3 + i) for i, b in enumerate(f)]
    pmask = synth.blob_mask(rows, cols, cover=0.3, sigma=4.0, seed=9, clear_border=False)
    pwant, _ = port.poisson_blend(f, g, pmask, tol=1e-13, max_it=10**6)
    pgot = sa.blend_images_poisson(f, g, pmask, tolerance=1e-12, max_iterations=10**6)
    for b in range(2):
        assert rel_max_abs(pgot[b], pwant[b], pmask) < 1e-7
    unchanged = sa.blend_images_poisson(f, g, pmask, tolerance=1e-13, max_iterations=2)  # poisson.cpp:263-269
    assert all(np.array_equal(u, a) for u, a in zip(unchanged, f))
    lab, k = core.find_connected_components(np.asfortranarray(pmask))
    wl, wk = port.label_components(pmask)
    assert k == wk and np.array_equal(lab, wl)
    with pytest.raises(NotImplementedError):
        sa.detect
