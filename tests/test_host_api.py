"""CPU: host-side mirror of the reference's Python surface -- dtype rules and error behaviour that are decided before
any device work (src/main.cpp:49-58, laplace.cpp:124-127, poisson.cpp:154-160)."""
from __future__ import annotations

import numpy as np
import pytest

import satellite_approximation_b200 as sab


def test_exports_match_reference_module():
    import os

    import satellite_approximation as sa  # the reference-named package: pybind11 _core when built, else the ctypes mirror

    for mod in (sab, sa):
        for name in ("LogLevel", "Path", "set_log_level", "filling_missing_portions_smooth_boundaries",
                     "blend_images_poisson"):  # fmt: skip
            assert hasattr(mod, name), (mod.__name__, name)
        assert os.fspath(mod.Path("a/b.tif")) == "a/b.tif"  # src/main.cpp:20-22: constructible from a str
    with pytest.raises(NotImplementedError):
        sa.detect  # cloud / shadow detection is outside the path
    assert [m.name for m in sab.LogLevel] == ["Debug", "Info", "Warn", "Error", "Critical"]  # src/main.cpp:24-29
    sab.set_log_level(sab.LogLevel.Warn)


def test_laplace_noconvert_dtype_rules():
    img = np.zeros((4, 5), np.float64)
    mask = np.zeros((4, 5), bool)
    with pytest.raises(TypeError):
        sab.filling_missing_portions_smooth_boundaries(img.astype(np.float32), mask)  # noconvert: no silent cast
    with pytest.raises(TypeError):
        sab.filling_missing_portions_smooth_boundaries(img, mask.astype(np.uint8))
    with pytest.raises(TypeError):
        sab.filling_missing_portions_smooth_boundaries(img[0], mask)


def test_laplace_size_mismatch_raises_runtime_error():
    with pytest.raises(RuntimeError):  # laplace.cpp:124-127
        sab.filling_missing_portions_smooth_boundaries(np.zeros((4, 5)), np.zeros((4, 6), bool))


def test_poisson_size_mismatch_returns_inputs_unchanged():
    f = [np.arange(20.0).reshape(4, 5)]
    g = [np.zeros((4, 6))]
    out = sab.blend_images_poisson(f, g, np.zeros((4, 5), bool))  # poisson.cpp:154-157: log and return
    assert len(out) == 1 and np.array_equal(out[0], f[0]) and out[0].flags.f_contiguous
    out = sab.blend_images_poisson(f, [np.zeros((4, 5))], np.zeros((5, 4), bool))
    assert np.array_equal(out[0], f[0])
    assert sab.blend_images_poisson([], [], np.zeros((0, 0), bool)) == []


def test_region_map_from_labels():
    lab = np.array([[1, 0, 2], [1, 0, 2], [0, 0, 2]], np.int32)
    cc = sab.ConnectedComponents(lab, 2)
    rm = cc.region_map
    assert set(rm) == {1, 2}
    assert rm[1].tolist() == [[0, 0], [1, 0]] and rm[2].tolist() == [[0, 2], [1, 2], [2, 2]]
    assert sab.ConnectedComponents(np.zeros((3, 3), np.int32), 0).region_map == {}


def test_stride_helpers():
    from satellite_approximation_b200 import _capi

    a = np.zeros((6, 8))
    assert _capi.is_dense_2d(a) and _capi.is_dense_2d(np.asfortranarray(a))
    assert _capi.is_dense_2d(a[:, :5]) and _capi.is_dense_2d(a[1:4])
    assert not _capi.is_dense_2d(a[::2, ::2])
    assert _capi.element_strides(np.asfortranarray(a)) == (1, 6)


def test_image_helpers_of_the_offset_demo():
    """read_image / image_list_to_cv arithmetic (approx/source/utils.cpp:16-60), valid_pixel (approx/utils.h:101-105) and
    highlight_area_replaced (poisson.cpp:305-321): host-side value helpers around the offset overload."""
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    ch = sab.image_to_channels(img)
    assert len(ch) == 3 and ch[0].dtype == np.float64
    assert np.allclose(ch[0], (img[..., 2] / 255.0) ** (1 / 2.2)) and np.allclose(ch[2], (img[..., 0] / 255.0) ** (1 / 2.2))
    back = sab.channels_to_image(ch)
    assert back.dtype == np.uint8 and np.max(np.abs(back.astype(int) - img.astype(int))) <= 1  # truncating cast
    assert sab.channels_to_image(ch[:2]) is None
    rep = [np.full((3, 4), 0.5) for _ in range(3)]
    for c in rep:
        c[0, :] = 1.0   # white key: truncates to 1 in all three channels
    rep[1][0, 1] = 0.99  # one channel off the key -> a valid pixel
    m = sab.valid_pixel_mask(rep)
    assert m.tolist() == [[False, True, False, False], [True] * 4, [True] * 4]
    big = [np.zeros((6, 8)) for _ in range(3)]
    sab.highlight_area_replaced(big, rep, 2, 3, (0.1, 0.2, 0.3))
    want = np.zeros((6, 8), bool)
    want[2:5, 3:7] = m
    for k, v in enumerate((0.1, 0.2, 0.3)):
        assert np.array_equal(big[k] == v, want)


def test_prepost_dtype_rules_before_any_device_work():
    img = np.zeros((4, 5, 3), np.uint8)
    with pytest.raises(TypeError):
        sab.Context.apply_laplace(None, img.astype(np.float64), img)
    with pytest.raises(TypeError):
        sab.Context.morph_close_mask(None, np.zeros((4, 5), np.float32))


def test_write_perf_info_csv(tmp_path):
    """PerfInfo::write (poisson.cpp:13-18): region_size,tolerance,max_iterations,iterations,error,solve_time appended."""
    recs = [sab.SolveStats(unknowns=633573, tolerance=1e-6, max_iterations=316786, iterations=989, error=9.96e-7,
                           solve_ms=12500.0),
            sab.SolveStats(unknowns=4, tolerance=0.5, max_iterations=2, iterations=1, error=0.0, solve_ms=0.25)]  # fmt: skip
    p = tmp_path / "perf_info.csv"
    sab.write_perf_info(p, recs[:1])
    sab.write_perf_info(p, recs[1:])  # appends
    assert p.read_text() == "633573,1e-06,316786,989,9.96e-07,12.5\n4,0.5,2,1,0,0.00025\n"  # operator<< on doubles = %g


def _core_or_skip():
    try:
        from satellite_approximation import _core
    except ImportError:
        pytest.skip("satellite_approximation._core is not built (make -C cpp pybind needs Eigen headers)")
    return _core


def test_cpp_shim_host_pieces_match_the_python_mirror(tmp_path):
    """The dependency-free host functions of the C++ `approx` shim (cpp/): highlight_area_replaced (poisson.cpp:305-321),
    the ranking rule of find_good_close_image (poisson.cpp:323-349), utils::Date day arithmetic and
    utils::find_directory_contents -- against the Python mirror and the datetime module."""
    import datetime as dt

    from satellite_approximation_b200 import scenes as sc

    core = _core_or_skip()
    rng = np.random.default_rng(2)
    ins = [rng.random((12, 15)) for _ in range(3)]
    rep = [rng.random((5, 6)) * 0.9 for _ in range(3)]
    for ch in rep:
        ch[1:3, 2:5] = 1.25  # the white key: truncates to 1 in all three channels
    want = [a.copy() for a in ins]
    sab.highlight_area_replaced(want, rep, 4, 7, (0.1, 0.2, 0.3))
    got = core.highlight_area_replaced(ins, rep, 4, 7, (0.1, 0.2, 0.3))
    assert all(np.array_equal(g, w) for g, w in zip(got, want)) and not np.array_equal(got[0], ins[0])
    # Date: days since the epoch, validation
    for s in ("1970-01-01", "2019-05-22", "2020-02-29", "2000-03-01", "1900-03-01", "2019-5-2", "2019/12/31"):
        assert core.date_days(s) == (sc.parse_simple_date(s) - dt.date(1970, 1, 1)).days, s
    for bad in ("2019-02-29", "2019-13-01", "2019-05", "2019-05-22x", "x"):
        with pytest.raises(IndexError):  # std::out_of_range
            core.date_days(bad)
    # find_good_close_image on the rows of test_scenes.test_find_good_close_image
    rows = [("2019-04-30", 0.0), ("2019-05-02", 0.05), ("2019-05-20", 0.90), ("2019-06-11", 0.10)]
    with sc.DataBase(tmp_path) as db:
        for d, p in rows + [("2019-05-22", 0.30)]:
            db.write_detection_result(d, True, True, 0, 0, p)
        for w in (0.0, 0.01, 0.05, 0.5, 1.0):
            assert core.find_good_close_image("2019-05-22", w, rows, 0.30) == sc.find_good_close_image("2019-05-22", w, db), w
    assert core.find_good_close_image("2019-05-22", 0.5, [], 0.3) == ""
    assert core.find_good_close_image("2019-05-21", 1.0, rows, float("nan")) == "2019-05-20"
    with pytest.raises(RuntimeError):
        core.find_good_close_image("2019-05-22", 1.5, rows, 0.3)
    # find_directory_contents
    (tmp_path / "2019-05-22").mkdir()
    (tmp_path / "2019-05-22" / "B04.tif").write_bytes(b"")
    (tmp_path / "2019-05-23").mkdir()
    for name in ("2019-05-22", "2019-05-23", "logs"):
        assert core.find_directory_contents(str(tmp_path / name)) == sc.find_directory_contents(tmp_path / name).value


def test_cpp_shim_behaviour_on_the_cpu(tmp_path):
    """The C++ `approx` shim behind satellite_approximation._core, exercised on the CPU: a libsatfill.so built from
    tests/fake_satfill.c (the eight C-ABI entry points the shim imports, answered by the ORACLE) is put in front of the
    real library through LD_LIBRARY_PATH in a child process.  What is checked is the shim's own host logic -- Eigen
    casters and layouts, copies, defaults, the reference's error behaviours -- not the device code (tests -m gpu)."""
    import os
    import subprocess
    import sys

    import oracle

    _core_or_skip()
    oracle.port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    odir = os.path.join(root, "oracle", "_build")
    fake = tmp_path / "fake"
    fake.mkdir()
    r = subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "fake_satfill.c"), "-o", str(fake / "libsatfill.so"), "-L", odir,
                        "-loracle", f"-Wl,-rpath,{odir}"], capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=str(fake) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "core_fake_worker.py")], env=env, capture_output=True,
                       text=True, timeout=600)  # fmt: skip
    assert r.returncode == 0 and "core-on-fake ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_cpp_api_gated_on_opencv_and_the_database(tmp_path):
    """cv::Mat apply_laplace(cv::Mat const&, cv::Mat const&, f64) (laplace.h:31) and find_good_close_image(std::string
    const&, f64, DataBase&) (poisson.h:63) with the reference's signatures: compiled against tests/fake_opencv (a minimal
    cv::Mat: this image has no OpenCV C++ headers) and a stand-in database, run on the CPU against the fake C-ABI, checked
    against the oracle's restatement of laplace.cpp:134-168 and the Python mirror of the picker."""
    import os
    import subprocess

    import oracle

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    eigen = os.environ.get("EIGEN_DIR", "/root/reference/thirdparty/eigen-master")
    shim = os.path.join(root, "satellite_approximation_b200", "lib", "libapprox_satfill.so")
    if not os.path.isdir(eigen) or not os.path.exists(shim):
        pytest.skip("needs Eigen headers and the built C++ shim (make -C cpp)")
    oracle.port()
    odir = os.path.join(root, "oracle", "_build")
    fake = tmp_path / "fake"
    fake.mkdir()
    r = subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "fake_satfill.c"), "-o", str(fake / "libsatfill.so"), "-L", odir,
                        "-loracle", f"-Wl,-rpath,{odir}"], capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr
    exe = tmp_path / "cpp_gated_api"
    r = subprocess.run(["g++", "-O1", "-std=c++20", "-Wall", "-I", os.path.join(root, "tests", "fake_opencv"), "-I",
                        os.path.join(root, "cpp", "include"), "-I", os.path.join(root, "include"), "-I", eigen,
                        os.path.join(root, "tests", "cpp_gated_api.cpp"), "-o", str(exe), "-L", os.path.dirname(shim),
                        "-lapprox_satfill", "-L", str(fake), "-lsatfill", f"-Wl,-rpath,{os.path.dirname(shim)}"],
                       capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr[-3000:]
    rows, cols = 37, 45
    env = dict(os.environ, LD_LIBRARY_PATH=str(fake) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), str(rows), str(cols), str(tmp_path)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    # read_image / image_list_to_cv / write_image (utils.cpp:16-68): size, worst level error of a write -> read round trip of
    # a ramp (the reference's truncating encode lands one level low wherever pow(pow(v, 1/2.2), 2.2) * 255 comes out a hair under
    # the integer: about a third of the levels), B G R order in the file, IOError, nothing written for 2 channels
    img = lines.pop().split()
    assert img[0] == "images" and [int(v) for v in img[1:3]] == [4, 256] and int(img[3]) <= 1 and int(img[4]) > 3 * 4 * 256 * 0.5
    assert [int(v) for v in img[5:]] == [255, 0, 255, 1, 0], img
    assert "less than 3 channels" in r.stderr
    assert lines[0] == f"image {rows} {cols}"
    vals = np.array([[float(x) for x in ln.split()] for ln in lines[1 : 1 + rows * cols * 3]])
    image = vals[:, 0].astype(np.uint8).reshape(rows, cols, 3)
    invalid = vals[:, 1].astype(np.uint8).reshape(rows, cols, 3)
    got = vals[:, 2].reshape(rows, cols, 3)
    want, mask = oracle.apply_laplace(image, invalid, 220.0, tol=1e-13)
    assert mask.any() and np.max(np.abs(got - want)) < 1e-6
    assert np.array_equal(got[~mask], image[~mask].astype(np.float64))
    # the picker: nearest date for weight 1, cleanest date for weight 0, the date itself when it is cleaner than its best
    # neighbour, "" (and only ONE query) without neighbours, and the weight check before any query
    assert lines[-1] == "picker 1 2019-05-20 2019-06-01 2019-05-22 [] 1 weight-error 0", lines[-1]


def test_find_good_close_image_cpp_and_python_agree_on_random_tables(tmp_path):
    """The ranking rule of find_good_close_image (poisson.cpp:323-349) in both host languages on random `dates` tables
    (hypothesis): same answer for every date, weight and invalid fraction, ties included."""
    import datetime as dt

    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from satellite_approximation_b200 import scenes as sc

    core = _core_or_skip()
    days = st.integers(0, 200).map(lambda k: dt.date(2019, 11, 1) + dt.timedelta(days=k))  # crosses a year boundary
    frac = st.sampled_from([0.0, 0.05, 0.1, 0.25, 0.25, 0.5, 0.9, 1.0])  # repeated values: ties
    counter = [0]

    @hyp.settings(max_examples=40, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(rows=st.dictionaries(days, frac, min_size=1, max_size=12), date=days, own=st.one_of(st.none(), frac),
               w=st.sampled_from([0.0, 0.01, 0.3, 0.5, 0.99, 1.0]))  # fmt: skip
    def run(rows, date, own, w):
        counter[0] += 1
        base = tmp_path / f"t{counter[0]}"
        base.mkdir()
        rows = dict(rows)
        rows.pop(date, None)
        with sc.DataBase(base) as db:
            for d, p in rows.items():
                db.write_detection_result(d.isoformat(), True, True, 0.0, 0.0, p)
            if own is not None:
                db.write_detection_result(date.isoformat(), True, True, 0.0, 0.0, own)
            want = sc.find_good_close_image(date.isoformat(), w, db)
            close = [(i.date.isoformat(), i.percent_invalid) for i in db.select_close_images(date.isoformat())]
        got = core.find_good_close_image(date.isoformat(), w, close, float("nan") if own is None else own)
        assert got == want, (rows, date, own, w)

    run()
